"""Developer tool (GPU box): band height sweep of the colour kernel on frame batches (config 3 shards)."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)

def timeit(I, p, q, r, iters, env):
    for k, v in env.items(): api.set_option(k, int(v))
    n, h, w = p.shape
    f = lambda: api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, sp)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): f()
    e1.record(s); torch.cuda.synchronize()
    for k in env: api.set_option(k, -1)
    return e0.elapsed_time(e1) / iters

g = torch.Generator(device="cuda").manual_seed(0)
for n in (32, 16, 64):
    I = torch.rand((n, 1080, 1920, 3), device="cuda", generator=g); p = torch.rand((n, 1080, 1920), device="cuda", generator=g)
    q = torch.empty_like(p)
    row = {"frames": n, "r": 16, "default_ms": round(timeit(I, p, q, 16, 8, {"GF_TAPE": 0}), 4)}
    for hb in (1080, 540, 360, 270, 216, 180, 154, 135, 120, 108, 90, 72, 60):
        row[f"hb{hb}"] = round(timeit(I, p, q, 16, 8, {"GF_TAPE": 0, "GF_C4_HB": hb}), 4)
    for we in (100, 120, 130, 140):
        row[f"tape_we{we}"] = round(timeit(I, p, q, 16, 8, {"GF_TAPE": 1, "GF_C4_EDGE_WEIGHT": we}), 4)
    for wps in (4, 5):
        row[f"wps{wps}"] = round(timeit(I, p, q, 16, 8, {"GF_TAPE": 0, "GF_C4_WARPS_PER_SM": wps}), 4)
    print(json.dumps(row), flush=True)
    del I, p, q
