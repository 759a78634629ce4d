"""Developer tool (ONE GPU): throughput of the strip kernel against resident warps per SM (GF_S8_EXTRA_SMEM lowers the
residency; the band plan follows the slot count), BASELINE configs[4] strip 32768 x 4096 r=16 and the 4K r=8 frame."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402

api = pkg.api()
s = torch.cuda.current_stream()
sp = ctypes.c_void_p(s.cuda_stream)


def timeit(W, H, r, opts, iters=8):
    g = torch.Generator(device="cuda").manual_seed(0)
    I = torch.rand((H, W), device="cuda", generator=g)
    p = torch.rand((H, W), device="cuda", generator=g)
    q = torch.empty_like(I)
    for k, v in opts.items():
        api.set_option(k, v)
    f = lambda: api.call("gf_guided_gray", I.data_ptr(), p.data_ptr(), q.data_ptr(), None, None, W, H, 0, 0, 0, 0, r, 1e-2, 0, sp)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        f()
    e1.record(s)
    torch.cuda.synchronize()
    for k in opts:
        api.set_option(k, -1)
    return e0.elapsed_time(e1) / iters, api.last_kernel()


# ring of gf_s8: (2r+1) rows x 8 x (32 - 4 ceil(r/8)) float2 -> r=8: 30.5 kB (7 warps/SM), r=16: 50.7 kB (4 warps/SM)
for (W, H, r, ring_kb, wmax) in ((32768, 4096, 16, 50.7, 4), (3840, 2160, 8, 30.5, 7), (7680, 4320, 8, 30.5, 7)):
    for warps in range(wmax, 0, -1):
        # per-CTA shared memory such that exactly `warps` CTAs fit: 228 kB / (ring + extra + 1 kB)
        extra = 0 if warps == wmax else int((228.0 / (warps + 0.5) - 1.0 - ring_kb) * 1024)
        opts = {"GF_WS": 0, "GF_S8_EXTRA_SMEM": max(0, extra)}
        ms, k = timeit(W, H, r, opts)
        print(json.dumps({"w": W, "h": H, "r": r, "kernel": k, "target_warps_per_sm": warps, "extra_smem": max(0, extra), "ms": round(ms, 4),
                          "gpix_s": round(W * H / ms / 1e6, 1)}), flush=True)
