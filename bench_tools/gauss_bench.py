"""Developer tool (GPU box): Gaussian blur timings, 8 B/px algorithmic."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for (w, h, r, sig) in [(3840, 2160, 1, 0.5), (3840, 2160, 4, 1.5), (3840, 2160, 8, 3.0), (3840, 2160, 16, 5.0), (7680, 4320, 8, 3.0), (1920, 1080, 8, 3.0)]:
    sets = [(torch.rand((h, w), device="cuda"), torch.empty((h, w), device="cuda")) for _ in range(6)]
    i = [0]
    def f():
        a, b = sets[i[0] % 6]; i[0] += 1
        api.call("gf_gaussian_gray", a.data_ptr(), b.data_ptr(), w, h, 0, 0, r, sig, sp)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(50): f()
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(json.dumps({"w": w, "h": h, "r": r, "sigma": sig, "us": round(ms * 1e3, 2), "gpix_s": round(w * h / ms / 1e6, 2),
                      "gbs_alg": round(8.0 * w * h / ms / 1e6, 1), "frac_of_6535": round(8.0 * w * h / ms / 1e6 / 6535.7, 3)}), flush=True)
    del sets
