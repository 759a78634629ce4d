"""Developer tool: the reference's own path-A demo shape -- 1080p, gray guide, 3-channel source, r=7, eps=0.3 (main.cpp:109-150)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
w, h = 1920, 1080
g = torch.Generator(device="cuda").manual_seed(0)
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for gch in (1, 3):
    sets = [(torch.rand((h, w) if gch == 1 else (h, w, 3), device="cuda", generator=g), torch.rand((h, w, 3), device="cuda", generator=g),
             torch.empty((h, w, 3), device="cuda")) for _ in range(8)]
    hnd = ctypes.c_void_p(); api.call("gf_create", ctypes.addressof(hnd), w, h, gch, 3)
    for env in ({}, {"GF_DISABLE_S8": "1"}):
        os.environ.update(env)
        f = lambda i: api.call("gf_run", hnd, sets[i % 8][0].data_ptr(), sets[i % 8][1].data_ptr(), sets[i % 8][2].data_ptr(), 7, 0.3, 1, 0, 0, 0, sp)
        for i in range(5): f(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(40): f(i)
        e1.record(s); torch.cuda.synchronize()
        print(f"class run ({gch},3) 1080p r=7 TRUNCATE", api.last_kernel(), round(e0.elapsed_time(e1) / 40 * 1e3, 1), "us", flush=True)
        for k in env: os.environ.pop(k)
    api.call("gf_destroy", hnd)
