"""Developer tool: the class API path (GuidedFilter::run = gf_run, TRUNCATE border) on a 4K gray frame."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
w, h = 3840, 2160
g = torch.Generator(device="cuda").manual_seed(0)
sets = [(torch.rand((h, w), device="cuda", generator=g), torch.rand((h, w), device="cuda", generator=g), torch.empty((h, w), device="cuda")) for _ in range(6)]
hnd = ctypes.c_void_p(); api.call("gf_create", ctypes.addressof(hnd), w, h, 1, 1)
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for env in ({}, {"GF_DISABLE_S8": "1"}):
    os.environ.update(env)
    f = lambda i: api.call("gf_run", hnd, sets[i % 6][0].data_ptr(), sets[i % 6][1].data_ptr(), sets[i % 6][2].data_ptr(), 8, 1e-2, 1, 0, 0, 0, sp)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(60): f(i)
    e1.record(s); torch.cuda.synchronize()
    print("class run 4K r=8 TRUNCATE", api.last_kernel(), round(e0.elapsed_time(e1) / 60 * 1e3, 1), "us", flush=True)
    for k in env: os.environ.pop(k)
