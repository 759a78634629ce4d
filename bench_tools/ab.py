"""A/B timing of differently compiled builds of libgf_b200 (developer tool).
    python bench_tools/ab.py lib1.so lib2.so ...     (paths relative to cudaimageprocessing_b200/)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "bench_tools"))
import sweep
out = []
for env in ({"GF_WP_WARPS_PER_SM": 12, "GF_WP_HB_MIN": 32}, {"GF_WP_WARPS_PER_SM": 16, "GF_WP_HB_MIN": 32}, {"GF_WP_WARPS_PER_SM": 10, "GF_WP_HB_MIN": 32}):
    out.append(sweep.time_gray(3840, 2160, 8, env=env))
out.append(sweep.time_gray(7680, 4320, 8, nsets=3, iters=20))
print("RESULT " + json.dumps([(o["us"], o["env"]) for o in out]))
''' % (ROOT, ROOT)

for lib in sys.argv[1:]:
    env = dict(os.environ)
    env["GF_LIB_PATH"] = os.path.join(ROOT, "cudaimageprocessing_b200", lib)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(lib, line[0][7:] if line else r.stderr[-500:], flush=True)
