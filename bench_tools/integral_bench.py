"""Developer tool: integral image (uint8 -> int32 SAT) timings; the reference reports 0.597 ms at 4K on sm_86."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for (w, h) in ((3840, 2160), (7680, 4320), (1920, 1080), (5910, 5941)):
    sets = [(torch.randint(0, 256, (h, w), dtype=torch.uint8, device="cuda"), torch.empty((h, w), dtype=torch.int32, device="cuda")) for _ in range(12)]
    scr = torch.empty((h // 16 + 2, w), dtype=torch.int32, device="cuda")
    f = lambda i: api.call("gf_integral_u8_i32", sets[i % 12][0].data_ptr(), sets[i % 12][1].data_ptr(), scr.data_ptr(), w, h, 0, 0, sp)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(60): f(i)
    e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    print(json.dumps({"w": w, "h": h, "us": round(us, 1), "gpix_s": round(w * h / us / 1e3, 1), "alg_gb_s_5Bpx": round(5.0 * w * h / us / 1e3, 1),
                      "moved_gb_s_13Bpx": round(13.0 * w * h / us / 1e3, 1)}), flush=True)
