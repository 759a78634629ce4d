"""The two multi-GPU workloads of BASELINE.json (configs[2] and configs[4]) as bench.py legs.

config 3  256 synthetic 1080p frames, RGB guide + 1-channel source, r=16, eps=1e-2: frames sharded over the ranks in
          contiguous blocks, no collective; STRONG scaling (the batch is fixed): ms per batch = max over ranks.
config 5  one 32768 x 32768 gray image, r=16, row strips.  Every rank keeps its strip in IPC-shared device buffers;
          one gf_run_strips call per rank and image pulls the 2r halo rows out of the neighbours' buffers (peer copies
          over NVLink) and runs the strip kernel.  Reported: the whole call, the kernel alone (second call on the
          already filled buffers without peers), and the seam check: the first and last `seam_rows` output rows of every
          rank against the C oracle (float64) on rows regenerated on the host -- pixels come from a counter-based hash
          of the GLOBAL (y, x), so any partition sees the same image.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np


def hash_rows_np(y0: int, y1: int, width: int, seed: int) -> np.ndarray:
    y = np.arange(y0, y1, dtype=np.uint64)[:, None]
    x = np.arange(width, dtype=np.uint64)[None, :]
    m = np.uint64(0xFFFFFFFF)
    v = (y * np.uint64(2654435761) + x * np.uint64(40503) + np.uint64(seed * 97)) & m
    v = ((v ^ (v >> np.uint64(15))) * np.uint64(2246822519)) & m
    v = ((v ^ (v >> np.uint64(13))) * np.uint64(3266489917)) & m
    v = v ^ (v >> np.uint64(16))
    return ((v >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def hash_rows_torch(torch, out, y0: int, width: int, seed: int):
    """Fills `out` (rows x width, cuda float32) with rows [y0, y0 + rows) of the synthetic image."""
    x = torch.arange(width, device=out.device, dtype=torch.int64)
    rows = out.shape[0]
    for c0 in range(0, rows, 1024):
        c1 = min(rows, c0 + 1024)
        y = torch.arange(y0 + c0, y0 + c1, device=out.device, dtype=torch.int64)[:, None]
        v = (y * 2654435761 + x * 40503 + seed * 97) & 0xFFFFFFFF
        v = ((v ^ (v >> 15)) * 2246822519) & 0xFFFFFFFF
        v = ((v ^ (v >> 13)) * 3266489917) & 0xFFFFFFFF
        v = v ^ (v >> 16)
        out[c0:c1] = (v >> 8).to(torch.float32) * (1.0 / 16777216.0)


def _max_over_ranks(torch, dist, v: float, world: int) -> float:
    if world == 1:
        return v
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config3(torch, dist, api, pkg, rank, world, barrier, frames=256, steps=3, r=16, eps=1e-2):
    from cudaimageprocessing_b200 import dist as D
    f0, f1 = D.shard_frames(frames, rank, world)
    n = f1 - f0
    stream = torch.cuda.current_stream()
    I = torch.empty((n, 1080, 1920, 3), device="cuda")
    p = torch.empty((n, 1080, 1920), device="cuda")
    g = torch.Generator(device="cuda")
    for k in range(n):                       # guide seed 100 + k, source seed 10000 + k (global frame index)
        g.manual_seed(100 + f0 + k)
        I[k].uniform_(generator=g)
        g.manual_seed(10000 + f0 + k)
        p[k].uniform_(generator=g)
    q = torch.empty_like(p)
    for _ in range(2):
        D.filter_frames(api, I, p, q, r, eps, 0, stream.cuda_stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        D.filter_frames(api, I, p, q, r, eps, 0, stream.cuda_stream)
    e1.record(stream)
    barrier()
    ms = _max_over_ranks(torch, dist, e0.elapsed_time(e1) / steps, world)
    kern = api.last_kernel()
    px = frames * 1080 * 1920
    out = {"workload": f"{frames} x 1920x1080 float32, RGB guide + 1-channel source, r={r}, eps={eps}, REFLECT101 (BASELINE configs[2])",
           "scaling": "strong", "n_gpus": world, "frames_per_gpu": n, "ms_per_batch": ms, "mpix_s": px / ms / 1e3,
           "alg_gb_s_per_gpu": 20.0 * px / world / ms / 1e6, "kernel": kern, "collective": "none (frames are independent)", "steps": steps}
    del I, p, q
    torch.cuda.empty_cache()
    return out


def config5(torch, dist, api, pkg, rank, world, barrier, size=32768, steps=3, r=16, eps=1e-2, seam_rows=64):
    from cudaimageprocessing_b200 import dist as D
    from oracle import c_oracle as C
    H = W = size
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    ps = D.PeerStrips(api, H, W, rank, world, r)
    y0, y1 = ps.y0, ps.y1
    hash_rows_torch(torch, ps.own_guide, y0, W, 7)
    hash_rows_torch(torch, ps.own_src, y0, W, 8)
    q = torch.empty((y1 - y0, W), device="cuda")
    for _ in range(2):
        barrier()
        ps.run(q, eps, 0, sp)
    t_all = t_k = 0.0
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    up, down = ps.up, ps.down
    for _ in range(steps):
        barrier()                                   # neighbours' rows are complete and stay put while halos are pulled
        e0.record(stream)
        ps.run(q, eps, 0, sp)                       # pull halos + strip kernel: ONE C call
        e1.record(stream)
        ps.up = ps.down = None
        ps.run(q, eps, 0, sp)                       # the kernel alone (halos are already in the buffers)
        e2.record(stream)
        ps.up, ps.down = up, down
        torch.cuda.synchronize()
        t_all += e0.elapsed_time(e1)
        t_k += e1.elapsed_time(e2)
    barrier()
    ms_all = _max_over_ranks(torch, dist, t_all / steps, world)
    ms_k = _max_over_ranks(torch, dist, t_k / steps, world)
    kern = api.last_kernel()
    # seam check against the oracle: rows next to this rank's upper and lower seam
    nt = len(__import__("os").sched_getaffinity(0))
    nt = max(1, nt // max(1, min(world, 8)))
    err = 0.0
    checked = 0
    t0 = time.perf_counter()
    for (a, b) in ((y0, min(y1, y0 + seam_rows)), (max(y0, y1 - seam_rows), y1)):
        lo, hi = max(0, a - 2 * r), min(H, b + 2 * r)
        # the oracle filters the block [lo, hi) as if it were an image: rows within 2r of an ARTIFICIAL cut are discarded
        lo2, hi2 = max(0, lo - 2 * r), min(H, hi + 2 * r)
        Ib, Pb = hash_rows_np(lo2, hi2, W, 7), hash_rows_np(lo2, hi2, W, 8)
        ref = C.guided_gray_f64(Ib, Pb, r, eps, 0, nt)[a - lo2:b - lo2]
        got = q[a - y0:b - y0].cpu().numpy()
        err = max(err, float(np.abs(got - ref).max()))
        checked += (b - a)
    t_check = time.perf_counter() - t0
    err = _max_over_ranks(torch, dist, err, world)
    px = H * W
    out = {"workload": f"{W}x{H} float32 gray, r={r}, eps={eps}, REFLECT101, row strips (BASELINE configs[4])", "scaling": "strong",
           "n_gpus": world, "rows_per_gpu": y1 - y0, "ms_total": ms_all, "ms_kernel": ms_k, "ms_halo_exchange": max(0.0, ms_all - ms_k),
           "mpix_s": px / ms_all / 1e3, "alg_gb_s_per_gpu": 12.0 * px / world / ms_k / 1e6,
           "halo_bytes_per_neighbour": 2 * r * W * 4 * 2, "kernel": kern,
           "exchange": "gf_run_strips: one kernel pulls the 2r halo rows out of the neighbours' IPC-mapped strip buffers over NVLink, the "
                       "strip kernel follows on the same stream; ranks are barrier-synchronised around the step. ms_halo_exchange = ms_total "
                       "- the strip kernel alone (pull + launch gap after an idle GPU; at 1 GPU, with no pull at all, it is the launch gap)",
           "seam_check": {"max_abs_err_vs_oracle_f64": err, "rows_per_rank": checked, "what": f"first and last {seam_rows} output rows of every "
                          "rank's strip (both sides of every seam) against oracle/gf_oracle.c in float64 on regenerated rows", "seconds": t_check},
           "steps": steps}
    ps.close()
    del q
    torch.cuda.empty_cache()
    return out
