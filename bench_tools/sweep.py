"""Timing sweeps on the GPU box (developer tool, not part of the product or the tests).
    python bench_tools/sweep.py [case ...]
"""
import ctypes
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402

api = pkg.api()


def time_gray(w, h, r, nsets=6, iters=60, border=0, env=None):
    for k, v in (env or {}).items():
        api.set_option(k, int(v))
    g = torch.Generator(device="cuda").manual_seed(0)
    sets = [(torch.rand((h, w), device="cuda", generator=g), torch.rand((h, w), device="cuda", generator=g),
             torch.empty((h, w), device="cuda")) for _ in range(nsets)]
    s = torch.cuda.current_stream()
    sp = ctypes.c_void_p(s.cuda_stream)

    def run(i):
        a, b, c = sets[i % nsets]
        api.call("gf_guided_gray", a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, w, h, 0, 0, 0, 0, r, 1e-2, border, sp)
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(iters):
        run(i)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    for k in (env or {}):
        api.set_option(k, -1)
    return {"w": w, "h": h, "r": r, "kernel": api.last_kernel(), "us": ms * 1e3, "gpix_s": w * h / ms / 1e6,
            "gbs_alg": 12.0 * w * h / ms / 1e6, "env": env or {}}


def main():
    if __name__ != "__main__":
        return
    out = []
    cases = sys.argv[1:] or ["4k"]
    if "4k" in cases:
        for wps, hbm in itertools.product([4, 8, 12, 16], [32, 48, 64, 96, 136]):
            out.append(time_gray(3840, 2160, 8, env={"GF_WP_WARPS_PER_SM": wps, "GF_WP_HB_MIN": hbm}))
            print(json.dumps(out[-1]), flush=True)
        out.append(time_gray(3840, 2160, 8, env={"GF_DISABLE_WP": 1}))
        print(json.dumps(out[-1]), flush=True)
        out.append(time_gray(3840, 2160, 8, env={"GF_DISABLE_FAST": 1}))
        print(json.dumps(out[-1]), flush=True)
    if "borders" in cases:
        for b in (0, 1, 2):
            for env in ({}, {"GF_WP_WARPS_PER_SM": 8}, {"GF_WP_WARPS_PER_SM": 24, "GF_WP_BIG": 0}, {"GF_WP_WARPS_PER_SM": 36, "GF_WP_BIG": 0}):
                o = time_gray(3840, 2160, 8, border=b, env=env)
                o["border"] = b
                out.append(o)
                print(json.dumps(out[-1]), flush=True)
    if "r16" in cases:
        for env in ({}, {"GF_WP_LARGE": 1}, {"GF_WP_LARGE": 1, "GF_WP_BIG": 1}):
            for (w, h, r) in [(3840, 2160, 16), (3840, 2160, 12), (7680, 4320, 16)]:
                out.append(time_gray(w, h, r, nsets=3 if w * h > 3e7 else 6, iters=20, env=env))
                print(json.dumps(out[-1]), flush=True)
    if "color" in cases:
        for (n, r) in [(8, 16), (8, 8), (32, 16)]:
            g = torch.Generator(device="cuda").manual_seed(0)
            I = torch.rand((n, 1080, 1920, 3), device="cuda", generator=g)
            p = torch.rand((n, 1080, 1920), device="cuda", generator=g)
            q = torch.empty_like(p)
            s_ = torch.cuda.current_stream(); sp = ctypes.c_void_p(s_.cuda_stream)
            def run():
                api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, 1920, 1080, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, sp)
            run(); run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_)
            for _ in range(3): run()
            e1.record(s_); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out.append({"case": "color_batch", "frames": n, "r": r, "kernel": api.last_kernel(), "ms": ms, "gpix_s": n * 1920 * 1080 / ms / 1e6, "gbs_alg": 20.0 * n * 1920 * 1080 / ms / 1e6})
            print(json.dumps(out[-1]), flush=True)
    if cases and cases[0] == "custom":      # custom W H R [iters]  (kernel env vars come from the shell)
        w, h, r = int(cases[1]), int(cases[2]), int(cases[3])
        it = int(cases[4]) if len(cases) > 4 else 6
        out.append(time_gray(w, h, r, nsets=2, iters=it))
        print(json.dumps(out[-1]), flush=True)
    if "one" in cases:
        out.append(time_gray(3840, 2160, 8, iters=6))
        print(json.dumps(out[-1]), flush=True)
    if "sizes" in cases:
        for (w, h, r) in [(1920, 1080, 8), (1920, 1080, 16), (3840, 2160, 4), (3840, 2160, 7), (3840, 2160, 16), (7680, 4320, 8),
                          (7680, 4320, 16), (7680, 4320, 32), (16384, 8192, 16)]:
            out.append(time_gray(w, h, r, nsets=3 if w * h > 3e7 else 6, iters=20))
            print(json.dumps(out[-1]), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
