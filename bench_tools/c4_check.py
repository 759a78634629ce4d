"""Developer tool (GPU box): tuned colour kernel vs the generic one + timings (config 3 frames)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()

def run(I, p, r, env=None):
    for k, v in (env or {}).items(): api.set_option(k, int(v))
    n, h, w = p.shape
    q = torch.full_like(p, float("nan"))
    api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, None)
    torch.cuda.synchronize()
    k = api.last_kernel()
    for kk in (env or {}): api.set_option(kk, -1)
    return q, k

def timeit(I, p, r, iters=5, env=None):
    for k, v in (env or {}).items(): api.set_option(k, int(v))
    n, h, w = p.shape
    q = torch.empty_like(p)
    s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
    f = lambda: api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, sp)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): f()
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    k = api.last_kernel()
    for kk in (env or {}): api.set_option(kk, -1)
    return {"frames": n, "w": w, "h": h, "r": r, "kernel": k, "ms": round(ms, 3), "gpix_s": round(n * w * h / ms / 1e6, 2),
            "gbs_alg": round(20.0 * n * w * h / ms / 1e6, 1), "env": env or {}}

g = torch.Generator(device="cuda").manual_seed(0)
for (n, h, w, r) in [(2, 1080, 1920, 16), (1, 540, 960, 8), (1, 300, 512, 4), (1, 400, 640, 12)]:
    I = torch.rand((n, h, w, 3), device="cuda", generator=g); p = torch.rand((n, h, w), device="cuda", generator=g)
    q1, k1 = run(I, p, r); q0, k0 = run(I, p, r, env={"GF_DISABLE_C4": 1})
    print(json.dumps({"case": "vs_generic", "shape": [n, h, w], "r": r, "k_new": k1, "k_old": k0, "max_diff": float((q1 - q0).abs().max()),
                      "nan": int(torch.isnan(q1).sum())}), flush=True)
I = torch.rand((32, 1080, 1920, 3), device="cuda", generator=g); p = torch.rand((32, 1080, 1920), device="cuda", generator=g)
for env in ({}, {"GF_C4_WARPS_PER_SM": 4}, {"GF_C4_HB": 270}, {"GF_C4_HB": 135}, {"GF_DISABLE_C4": 1}):
    print(json.dumps(timeit(I, p, 16, env=env)), flush=True)
print(json.dumps(timeit(I, p, 8)), flush=True)
print(json.dumps(timeit(I[:1], p[:1], 16, iters=20)), flush=True)
