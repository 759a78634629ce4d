"""Developer tool (GPU box): the s8 kernel against the older kernels and the KAT, plus timings.
    python bench_tools/s8_check.py [quick|full]
Honours GF_LIB_PATH (a variant build).  Prints JSON lines; appends to gpurun_out/s8_check.jsonl."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import cudaimageprocessing_b200 as pkg  # noqa: E402
import sweep  # noqa: E402

api = pkg.api()
OUT = os.path.join(ROOT, "gpurun_out", "s8_check.jsonl")
TAG = os.path.basename(os.environ.get("GF_LIB_PATH", "default"))


def emit(o):
    o["lib"] = TAG
    print(json.dumps(o), flush=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(o) + "\n")


def run_gray(I, p, r, eps, border, env=None):
    for k, v in (env or {}).items():
        api.set_option(k, int(v))
    h, w = I.shape
    q = torch.full_like(I, float("nan"))
    api.call("gf_guided_gray", I.data_ptr(), p.data_ptr(), q.data_ptr(), None, None, w, h, 0, 0, 0, 0, r, eps, border, None)
    torch.cuda.synchronize()
    k = api.last_kernel()
    for kk in (env or {}):
        api.set_option(kk, -1)
    return q, k


def compare(w, h, r, border, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    I = torch.rand((h, w), device="cuda", generator=g)
    p = torch.rand((h, w), device="cuda", generator=g)
    q1, k1 = run_gray(I, p, r, 1e-2, border)
    q0, k0 = run_gray(I, p, r, 1e-2, border, env={"GF_DISABLE_S8": 1})
    d = (q1 - q0).abs()
    emit({"case": "vs_old", "w": w, "h": h, "r": r, "border": border, "k_new": k1, "k_old": k0,
          "max_diff": float(d.max()), "nan": int(torch.isnan(q1).sum())})


def kat():
    from conftest import load_kat_full
    from oracle import gf_oracle as O
    k = load_kat_full()
    I, P = torch.from_numpy(k["I"]).cuda(), torch.from_numpy(k["P"]).cuda()
    for env in ({}, {"GF_S8_HB": 128}, {"GF_S8_HB": 512}, {"GF_DISABLE_S8": 1}):
        q, kn = run_gray(I, P, k["r"], k["eps"], 0, env=env)
        d = O.to_u8(q.cpu().numpy()).astype(int) - k["gold"].astype(int)
        emit({"case": "kat", "kernel": kn, "env": env, "flips": int(np.count_nonzero(d)), "max_lsb": int(np.abs(d).max())})


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    for (w, h, r, b) in [(3840, 2160, 8, 0), (3840, 2160, 8, 2), (1000, 700, 8, 0), (1920, 1080, 16, 0), (3840, 2160, 7, 0),
                         (2048, 1024, 4, 2)]:
        compare(w, h, r, b)
    kat()
    for wps in (4, 5, 6, 7):
        emit(sweep.time_gray(3840, 2160, 8, env={"GF_S8_WARPS_PER_SM": wps}))
    for hb in (40, 64, 96, 128):
        emit(sweep.time_gray(3840, 2160, 8, env={"GF_S8_HB": hb}))
    emit(sweep.time_gray(7680, 4320, 8, nsets=3, iters=20))
    emit(sweep.time_gray(16384, 8192, 8, nsets=2, iters=10))
    if mode == "full":
        compare(7680, 4320, 32, 0)
        for (w, h, r) in [(1920, 1080, 8), (3840, 2160, 16), (7680, 4320, 16), (3840, 2160, 4), (3840, 2160, 7), (7680, 4320, 32)]:
            emit(sweep.time_gray(w, h, r, nsets=3 if w * h > 3e7 else 6, iters=20))
            emit(sweep.time_gray(w, h, r, nsets=3 if w * h > 3e7 else 6, iters=20, env={"GF_DISABLE_S8": 1}))


if __name__ == "__main__":
    main()
