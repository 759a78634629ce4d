"""Small shapes through every strip mode of the s8 / c4 kernels, for compute-sanitizer memcheck:
    compute-sanitizer --tool memcheck python bench_tools/sanitize_cases.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
api = pkg.api()
g = torch.Generator(device="cuda").manual_seed(0)
n = 0
for (h, w) in [(70, 64), (70, 72), (90, 256), (90, 264), (150, 480), (140, 1000), (135, 1004)]:
    for r in (1, 4, 7, 8, 16, 32):
        if h < 4 * r + 2:
            continue
        for border in (0, 1, 2):
            s = (w + 7) // 8 * 8
            I = torch.rand((h, s), device="cuda", generator=g); p = torch.rand((h, s), device="cuda", generator=g)
            q = torch.empty((h, s), device="cuda")
            api.call("gf_guided_gray", I.data_ptr(), p.data_ptr(), q.data_ptr(), None, None, w, h, s, s, s, 0, r, 1e-2, border, None)
            torch.cuda.synchronize(); n += 1
    # strips: 3 row strips with halos
    for r in (4, 8):
        if w < 64: continue
        I = torch.rand((h, w if w % 8 == 0 else (w + 7) // 8 * 8), device="cuda", generator=g); p = torch.rand_like(I)
        s = I.shape[1]
        for k in range(3):
            y0, y1 = h * k // 3, h * (k + 1) // 3
            b0, b1 = max(0, y0 - 2 * r), min(h, y1 + 2 * r)
            q = torch.empty((y1 - y0, s), device="cuda")
            api.call("gf_guided_gray_strip", I[b0:].data_ptr(), p[b0:].data_ptr(), q.data_ptr(), w, h, b0, b1 - b0, y0, y1 - y0, s, s, s, r, 1e-2, 0, None)
            torch.cuda.synchronize(); n += 1
for (h, w) in [(70, 128), (70, 132), (80, 256), (90, 388)]:
    for r in (4, 8, 12, 16):
        if h < 4 * r + 2: continue
        I = torch.rand((2, h, w, 3), device="cuda", generator=g); p = torch.rand((2, h, w), device="cuda", generator=g); q = torch.empty_like(p)
        api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), 2, w, h, 3, 0, 0, 0, 0, 0, 0, r, 1e-2, 0, None)
        torch.cuda.synchronize(); n += 1
print("cases run:", n, "last kernel", api.last_kernel())
