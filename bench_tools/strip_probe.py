"""Developer tool (GPU box): one rank's strip job of the giga-image (config 5) on a single GPU, for several strip heights."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
W = H = 32768; r = 16
for rows, y0 in ((4096, 4096), (8192, 8192), (8192, 0), (8192, 24576), (8190, 8192), (16384, 16384), (10923, 8192)):
    b0, b1 = max(0, y0 - 2 * r), min(H, y0 + rows + 2 * r)
    I = torch.rand((b1 - b0, W), device="cuda"); p = torch.rand((b1 - b0, W), device="cuda"); q = torch.empty((rows, W), device="cuda")
    for env in ({}, {"GF_S8_EDGE_PCT": 100}):
        for k, v in env.items(): api.set_option(k, int(v))
        f = lambda: api.call("gf_guided_gray_strip", I.data_ptr(), p.data_ptr(), q.data_ptr(), W, H, b0, b1 - b0, y0, rows, 0, 0, 0, r, 1e-2, 0, sp)
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): f()
        e1.record(s); torch.cuda.synchronize()
        for k in env: api.set_option(k, -1)
        print(json.dumps({"rows": rows, "y0": y0, "env": env, "ms": round(e0.elapsed_time(e1) / 5, 4), "kernel": api.last_kernel(),
                          "us_per_1k_rows": round(e0.elapsed_time(e1) / 5 / rows * 1e6, 1)}), flush=True)
    del I, p, q
