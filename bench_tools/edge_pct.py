"""Developer tool: band height of the edge strips (GF_S8_EDGE_PCT) vs kernel time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import sweep
cases = [(3840, 2160, 16), (3840, 2160, 4), (3840, 2160, 7), (7680, 4320, 32), (7680, 4320, 16)]
for (w, h, r) in cases:
    for pct in (100, 85, 70, 62, 55, 48):
        o = sweep.time_gray(w, h, r, nsets=6 if w < 7000 else 3, iters=40 if w < 7000 else 12, env={"GF_S8_EDGE_PCT": pct})
        print(w, h, r, "pct", pct, round(o["us"], 1), "us", flush=True)
