"""Developer tool (GPU box): the tuned colour kernel under the three border rules, beside the generic kernel.

    python bench_tools/c4_borders.py > gpurun_out/c4_borders.jsonl

32 x 1080p RGB guide, r = 16 (BASELINE configs[2] per GPU at 8 GPUs) and one 1080p frame through the class API's border
(TRUNCATE).  CUDA events on the launching stream; inputs (32 frames = 1.06 GB) are larger than L2."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402

api = pkg.api()


def timeit(I, p, r, border, iters, generic=False):
    n, h, w = p.shape
    q = torch.empty_like(p)
    s = torch.cuda.current_stream()
    sp = ctypes.c_void_p(s.cuda_stream)
    if generic:
        api.set_option("GF_DISABLE_C4", 1)
    f = lambda: api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, border, sp)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        f()
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    k = api.last_kernel()
    if generic:
        api.set_option("GF_DISABLE_C4", -1)
    return {"frames": n, "w": w, "h": h, "r": r, "border": border, "kernel": k, "ms": round(ms, 3),
            "gpix_s": round(n * w * h / ms / 1e6, 2), "gbs_alg": round(20.0 * n * w * h / ms / 1e6, 1)}, q


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    I = torch.rand((32, 1080, 1920, 3), device="cuda", generator=g)
    p = torch.rand((32, 1080, 1920), device="cuda", generator=g)
    for border in (0, 2, 1):
        rec, q1 = timeit(I, p, 16, border, 10)
        rec0, q0 = timeit(I[:4], p[:4], 16, border, 2, generic=True)
        rec["generic_ms_per_frame"] = round(rec0["ms"] / 4, 3)
        rec["max_diff_vs_generic"] = float((q1[:4] - q0).abs().max())
        print(json.dumps(rec), flush=True)
    for border in (0, 2, 1):
        rec, _ = timeit(I[:1], p[:1], 16, border, 40)
        print(json.dumps(rec), flush=True)
        rec, _ = timeit(I[:1], p[:1], 8, border, 40)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
