"""Developer tool: gf_guided_gray_host on PAGEABLE (malloc'd) buffers -- staged through pinned planes by the library's copy
threads (default) against the driver's own pageable cudaMemcpyAsync (GF_HOST_STAGED=0), over threads and bands."""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
W, H = 3840, 2160
I = np.random.default_rng(0).random((H, W), dtype=np.float32)
p = np.random.default_rng(1).random((H, W), dtype=np.float32)
q = np.empty_like(I)
f = lambda: api.call("gf_guided_gray_host", I.ctypes.data, p.ctypes.data, q.ctypes.data, W, H, 8, 1e-2, 0)
api.set_option("GF_HOST_STAGED", 0)
f(); ref = q.copy()
settings = [{"GF_HOST_STAGED": 0}] + [{"GF_HOST_STAGED": 1, "GF_HOST_COPY_THREADS": t, "GF_HOST_STAGED_BANDS": b} for t in (4, 8, 12, 16) for b in (8, 12)]
for opts in settings:
    for k, v in opts.items(): api.set_option(k, v)
    q[:] = 0
    for _ in range(3): f()
    ok = float(np.abs(q - ref).max())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    for k in opts: api.set_option(k, -1)
    print(json.dumps({"opts": opts, "ms": round(ms, 3), "max_abs_diff_vs_driver_path": ok}), flush=True)
