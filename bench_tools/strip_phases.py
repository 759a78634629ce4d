"""Developer tool (ONE GPU): the parts of gf_run_strips on a BASELINE configs[4] strip (32768 x 4096, r=16) with both
neighbours faked by buffers on the same GPU -- what the main job, the pull and the seam jobs cost alone and together.

    python bench_tools/strip_phases.py > gpurun_out/strip_phases.jsonl

The "overlap_*_only" / "no_pull" settings need a library built with -DGF_STRIP_DEBUG (build.build_variant("libgf_strip_debug.so",
["-DGF_STRIP_DEBUG"]) and GF_LIB_PATH): the product library ignores GF_STRIP_DEBUG_SKIP, because skipping parts gives
wrong pixels."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402
from cudaimageprocessing_b200._capi import StripPeer  # noqa: E402

api = pkg.api()
W, ROWS, R, H = 32768, 4096, 16, 3 * 4096
g = torch.Generator(device="cuda").manual_seed(0)
own = [torch.rand((ROWS + 4 * R, W), device="cuda", generator=g) for _ in range(2)]
nb = [torch.rand((ROWS, W), device="cuda", generator=g) for _ in range(2)]      # one buffer pair stands for both neighbours
q = torch.empty((ROWS, W), device="cuda")
peer = StripPeer(nb[0].data_ptr(), nb[1].data_ptr(), W, W, 0, ROWS)
use = torch.cuda.Stream() if "--stream" in sys.argv else torch.cuda.current_stream()      # --stream: a blocking non-default stream
sp = ctypes.c_void_p(use.cuda_stream)


def run():
    api.call("gf_run_strips", own[0].data_ptr(), own[1].data_ptr(), q.data_ptr(), W, H, ROWS, ROWS, W, W, W, R, 1e-2, 0,
             ctypes.byref(peer), ctypes.byref(peer), sp)


def timeit(opts, iters=10):
    for k, v in opts.items():
        api.set_option(k, v)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    single = []
    for _ in range(3):                      # one call after an idle GPU (what the scaling bench times)
        e0.record(use)
        run()
        e1.record(use)
        torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1))
    e0.record(use)
    for _ in range(iters):
        run()
    e1.record(use)
    torch.cuda.synchronize()
    for k in opts:
        api.set_option(k, -1)
    return e0.elapsed_time(e1) / iters, min(single)


settings = [("pull_then_launch", {"GF_STRIP_OVERLAP": 0}),
            ("overlap_all", {"GF_STRIP_OVERLAP": 1}),
            ("overlap_no_pull", {"GF_STRIP_OVERLAP": 1, "GF_STRIP_DEBUG_SKIP": 2}),
            ("overlap_main_only", {"GF_STRIP_OVERLAP": 1, "GF_STRIP_DEBUG_SKIP": 6}),
            ("overlap_seams_only", {"GF_STRIP_OVERLAP": 1, "GF_STRIP_DEBUG_SKIP": 3}),
            ("overlap_pull_only", {"GF_STRIP_OVERLAP": 1, "GF_STRIP_DEBUG_SKIP": 5}),
            ("overlap_main_and_seams", {"GF_STRIP_OVERLAP": 1, "GF_STRIP_DEBUG_SKIP": 2})]
for name, opts in settings:
    ms, one = timeit(opts)
    print(json.dumps({"setting": name, "stream": "side" if "--stream" in sys.argv else "default", "ms_back_to_back": round(ms, 4),
                      "ms_single_call": round(one, 4), "kernel": api.last_kernel()}), flush=True)
