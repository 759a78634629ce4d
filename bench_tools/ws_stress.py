"""Race hunt for the warp-specialised kernel (developer tool): repeats one launch many times and compares EVERY result
with the float64 torch reference.   python bench_tools/ws_stress.py W H R [iters]   (GF_WS=1 GF_WS_K=.. in the env)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cudaimageprocessing_b200 as pkg  # noqa: E402
from bench_tools.ws_bench import ref_f64  # noqa: E402


def main():
    w, h, r = (int(x) for x in sys.argv[1:4])
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    api = pkg.api()
    g = torch.Generator(device="cuda").manual_seed(1)
    I = torch.rand((h, w), device="cuda", generator=g)
    p = torch.rand((h, w), device="cuda", generator=g)
    ref = ref_f64(I, p, r, 1e-2)
    q = torch.empty_like(I)
    junk = torch.empty((64 << 20,), device="cuda")      # perturbs timing between launches
    worst, bad = 0.0, 0
    for i in range(iters):
        q.fill_(float("nan"))
        if i % 3 == 0:
            junk.normal_()
        api.call("gf_guided_gray", I.data_ptr(), p.data_ptr(), q.data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, 0, None)
        torch.cuda.synchronize()
        d = (q.double() - ref).abs()
        e = float(torch.nan_to_num(d, nan=1e9).max())
        worst = max(worst, e)
        bad += e > 1e-4
    print(json.dumps({"w": w, "h": h, "r": r, "kernel": api.last_kernel(), "iters": iters, "worst_err": worst, "bad_launches": int(bad),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("GF_WS")}}))


if __name__ == "__main__":
    main()
