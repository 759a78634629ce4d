"""Developer tool: the e2e host call (gf_guided_gray_host, pinned buffers) against the number of pipeline bands and
their taper (band b is taper % as tall as band b-1)."""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
W, H = 3840, 2160
n = W * H * 4
ptrs = []
for _ in range(3):
    p = ctypes.c_void_p(); api.call("gf_host_alloc", ctypes.addressof(p), n); ptrs.append(p)
for i in range(2):
    np.ctypeslib.as_array(ctypes.cast(ptrs[i], ctypes.POINTER(ctypes.c_float)), shape=(H, W))[:] = np.random.default_rng(i).random((H, W), dtype=np.float32)
f = lambda: api.call("gf_guided_gray_host", ptrs[0], ptrs[1], ptrs[2], W, H, 8, 1e-2, 0)
for rep in range(2):
    for nb in (1, 2, 3, 4, 5, 6, 8):
        for taper in ((100,) if nb == 1 else (100, 80, 65, 55, 45, 35, 25)):
            api.set_option("GF_HOST_BANDS", nb); api.set_option("GF_HOST_TAPER_PCT", taper)
            for _ in range(3): f()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(20): f()
            torch.cuda.synchronize()
            print(json.dumps({"bands": nb, "taper_pct": taper, "ms": round((time.perf_counter() - t0) / 20 * 1e3, 3), "rep": rep}), flush=True)
