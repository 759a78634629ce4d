"""Developer tool: PCIe floor for the e2e leg (4K gray: 66.4 MB up, 33.2 MB down, pinned)."""
import time, torch
n = 3840 * 2160
a = torch.empty(2 * n, dtype=torch.float32).pin_memory(); b = torch.empty(n, dtype=torch.float32).pin_memory()
da = torch.empty(2 * n, dtype=torch.float32, device="cuda"); db = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
def up():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
def down():
    with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
def both(): up(); down()
print({"h2d_66MB_ms": round(t(up), 3), "d2h_33MB_ms": round(t(down), 3), "both_concurrent_ms": round(t(both), 3)})
