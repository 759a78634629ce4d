"""N1: the scan path (float64 row prefixes + column pass, gf_scan.cuh) timed against the fused
sliding-window kernels on the same frames -- BASELINE configs[3] (8K gray r=32) and the headline shape.
    python bench_tools/scan_vs_sliding.py > profiles/r2_scan_vs_sliding.jsonl"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cudaimageprocessing_b200 as pkg  # noqa: E402


def main():
    api = pkg.api()
    s = torch.cuda.current_stream()
    sp = ctypes.c_void_p(s.cuda_stream)
    for (w, h, r) in ((7680, 4320, 32), (7680, 4320, 8), (3840, 2160, 8), (3840, 2160, 16), (3840, 2160, 32), (3840, 2160, 64)):
        g = torch.Generator(device="cuda").manual_seed(0)
        nsets = 3
        sets = [(torch.rand((h, w), device="cuda", generator=g), torch.rand((h, w), device="cuda", generator=g),
                 torch.empty((h, w), device="cuda")) for _ in range(nsets)]
        res = {}
        outs = {}
        for name, scan in (("sliding", 0), ("scan", 1)):
            api.set_option("GF_SCAN", scan)

            def run(i):
                a, b, c = sets[i % nsets]
                api.call("gf_guided_gray", a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, 0, sp)
            for i in range(nsets):
                run(i)
            torch.cuda.synchronize()
            outs[name] = sets[0][2].clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20 if scan == 0 else 6
            e0.record(s)
            for i in range(n):
                run(i)
            e1.record(s)
            torch.cuda.synchronize()
            res[name] = {"kernel": api.last_kernel(), "us": e0.elapsed_time(e1) / n * 1e3}
        api.set_option("GF_SCAN", -1)
        diff = float((outs["scan"] - outs["sliding"]).abs().max())
        print(json.dumps({"w": w, "h": h, "r": r, "sliding": res["sliding"], "scan": res["scan"],
                          "scan_over_sliding": res["scan"]["us"] / res["sliding"]["us"], "max_abs_diff": diff,
                          "scan_bytes_per_px_model": 6 * 32 + 24, "sliding_bytes_per_px": 12}), flush=True)
        del sets, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
