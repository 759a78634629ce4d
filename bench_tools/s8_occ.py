"""Developer tool: throughput vs resident warps per SM (8K, fixed band height)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import sweep
ring = 30464
for w in (1, 2, 3, 4, 5, 6, 7):
    extra = (228 * 1024) // w - 1024 - ring - 64 if w < 7 else 0
    o = sweep.time_gray(7680, 4320, 8, nsets=3, iters=10, env={"GF_S8_EXTRA_SMEM": max(0, extra), "GF_S8_HB": 135})
    print(w, round(o["us"], 1), round(o["gpix_s"], 1), flush=True)
