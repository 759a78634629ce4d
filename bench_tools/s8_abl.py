"""Developer tool: timing of ablated builds (GF_S8_ABL bits) -- where does the iteration time go?"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, json, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "bench_tools"))
import sweep
out = {"us_4k": round(sweep.time_gray(3840, 2160, 8)["us"], 1), "us_8k": round(sweep.time_gray(7680, 4320, 8, nsets=3, iters=20)["us"], 1),
       "us_8k_w4": round(sweep.time_gray(7680, 4320, 8, nsets=3, iters=20, env={"GF_S8_EXTRA_SMEM": 26000, "GF_S8_HB": 135})["us"], 1)}
print("RESULT " + json.dumps(out))
''' % (ROOT, ROOT)
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["GF_LIB_PATH"] = os.path.join(ROOT, "cudaimageprocessing_b200", lib)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(lib, line[0][7:] if line else r.stderr[-800:], flush=True)
