"""Developer tool (ONE GPU): the equal-cost "tape" split (GF_TAPE=1) against uniform bands for the r = 16 gray kernel on the
strip geometry of BASELINE configs[4] (32768 x 4096 per GPU at 8 GPUs), the whole 32768^2 image, and 4K / 8K frames.
Result (B200): the tape does not help gray strips (0.928 vs 0.923 ms) and costs 7 % on the whole image."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
def timeit(W, H, r, opts, iters=6):
    g = torch.Generator(device="cuda").manual_seed(0)
    I = torch.rand((H, W), device="cuda", generator=g); p = torch.rand((H, W), device="cuda", generator=g); q = torch.empty_like(I)
    for k, v in opts.items(): api.set_option(k, v)
    f = lambda: api.call("gf_guided_gray", I.data_ptr(), p.data_ptr(), q.data_ptr(), None, None, W, H, 0, 0, 0, 0, r, 1e-2, 0, sp)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): f()
    e1.record(s); torch.cuda.synchronize()
    for k in opts: api.set_option(k, -1)
    del I, p, q
    return e0.elapsed_time(e1) / iters
for (W, H, r) in ((32768, 4096, 16), (32768, 32768, 16), (3840, 2160, 16), (7680, 4320, 16)):
    for opts in ({"GF_WS": 0}, {"GF_WS": 0, "GF_TAPE": 1}, {"GF_WS": 0, "GF_TAPE": 1, "GF_S8_EDGE_PCT": 100}):
        print(json.dumps({"w": W, "h": H, "r": r, "opts": opts, "ms": round(timeit(W, H, r, opts), 4)}), flush=True)
