"""Developer tool (GPU box): KAT flips + 4K/8K timings for every variant library given.
    python bench_tools/s8_variants.py libgf_v_a.so libgf_v_b.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "bench_tools")); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
import sweep, s8_check
from conftest import load_kat_full
from oracle import gf_oracle as O
k = load_kat_full()
I, P = torch.from_numpy(k["I"]).cuda(), torch.from_numpy(k["P"]).cuda()
q, kn = s8_check.run_gray(I, P, k["r"], k["eps"], 0)
d = O.to_u8(q.cpu().numpy()).astype(int) - k["gold"].astype(int)
out = {"kat_flips": int(np.count_nonzero(d)), "kernel": kn}
out["us_4k"] = sweep.time_gray(3840, 2160, 8)["us"]
out["us_8k"] = sweep.time_gray(7680, 4320, 8, nsets=3, iters=20)["us"]
print("RESULT " + json.dumps(out))
''' % (ROOT, ROOT, ROOT)

for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["GF_LIB_PATH"] = os.path.join(ROOT, "cudaimageprocessing_b200", lib)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(lib, line[0][7:] if line else r.stderr[-800:], flush=True)
