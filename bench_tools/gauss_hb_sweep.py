"""Developer tool (GPU box): band height sweep of the 4-columns-per-thread Gaussian kernel."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
s = torch.cuda.current_stream(); sp = ctypes.c_void_p(s.cuda_stream)
for (w, h, r, sig) in [(3840, 2160, 8, 3.0), (3840, 2160, 4, 1.5), (3840, 2160, 1, 0.5), (1920, 1080, 8, 3.0), (7680, 4320, 8, 3.0)]:
    sets = [(torch.rand((h, w), device="cuda"), torch.empty((h, w), device="cuda")) for _ in range(6)]
    row = {"w": w, "h": h, "r": r}
    for hb in (0, 16, 24, 32, 40, 48, 64, 72, 96, 128, 192, 270):
        if hb: api.set_option("GF_GAUSS_HB", int(hb))
        i = [0]
        def f():
            a, b = sets[i[0] % 6]; i[0] += 1
            api.call("gf_gaussian_gray", a.data_ptr(), b.data_ptr(), w, h, 0, 0, r, sig, sp)
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(40): f()
        e1.record(s); torch.cuda.synchronize()
        api.set_option("GF_GAUSS_HB", -1)
        row["default" if not hb else f"hb{hb}"] = round(e0.elapsed_time(e1) / 40 * 1e3, 2)
    print(json.dumps(row), flush=True)
    del sets
