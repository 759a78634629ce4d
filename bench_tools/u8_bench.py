"""Developer tool: uint8 in/out guided filter (fused conversions) vs the float32 call on the same frame size."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import cudaimageprocessing_b200 as pkg
import sweep
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for (w, h) in ((3840, 2160), (7680, 4320)):
    n = 24 if w < 7000 else 8
    sets = [(torch.randint(0, 256, (h, w), dtype=torch.uint8, device="cuda"), torch.randint(0, 256, (h, w), dtype=torch.uint8, device="cuda"),
             torch.empty((h, w), dtype=torch.uint8, device="cuda")) for _ in range(n)]
    f = lambda i: api.call("gf_guided_gray_u8", sets[i % n][0].data_ptr(), sets[i % n][1].data_ptr(), sets[i % n][2].data_ptr(), w, h, 0, 0, 0, 8, 1e-2, 0, sp)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(60): f(i)
    e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    fl = sweep.time_gray(w, h, 8, nsets=6 if w < 7000 else 3, iters=30)
    print(json.dumps({"w": w, "h": h, "kernel": api.last_kernel(), "u8_us": round(us, 1), "u8_gpix_s": round(w * h / us / 1e3, 1),
                      "u8_alg_gb_s_3Bpx": round(3.0 * w * h / us / 1e3, 1), "f32_us": round(fl["us"], 1)}), flush=True)
