"""Developer tool: what a chunked H2D -> (kernel) -> D2H pipeline can reach on this box (pure torch copies)."""
import time, torch
W, H = 3840, 2160
hI = torch.empty((H, W), dtype=torch.float32).pin_memory(); hP = torch.empty((H, W), dtype=torch.float32).pin_memory()
hQ = torch.empty((H, W), dtype=torch.float32).pin_memory()
dI = torch.empty((H, W), device="cuda"); dP = torch.empty((H, W), device="cuda"); dQ = torch.empty((H, W), device="cuda")
up, comp, down = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(nb, interleave):
    evu = [torch.cuda.Event() for _ in range(nb)]; evk = [torch.cuda.Event() for _ in range(nb)]
    for b in range(nb):
        y0, y1 = H * b // nb, H * (b + 1) // nb
        with torch.cuda.stream(up):
            if interleave:
                dI[y0:y1].copy_(hI[y0:y1], non_blocking=True); dP[y0:y1].copy_(hP[y0:y1], non_blocking=True)
            else:
                dI[y0:y1].copy_(hI[y0:y1], non_blocking=True)
                dP[y0:y1].copy_(hP[y0:y1], non_blocking=True)
            evu[b].record(up)
        with torch.cuda.stream(comp):
            comp.wait_event(evu[b]); dQ[y0:y1].copy_(dI[y0:y1], non_blocking=True); evk[b].record(comp)
        with torch.cuda.stream(down):
            down.wait_event(evk[b]); hQ[y0:y1].copy_(dQ[y0:y1], non_blocking=True)
    down.synchronize(); comp.synchronize()
for nb in (1, 2, 4, 8, 16, 32):
    for _ in range(3): run(nb, True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): run(nb, True)
    torch.cuda.synchronize()
    print("bands", nb, round((time.perf_counter() - t0) / 20 * 1e3, 3), "ms", flush=True)
