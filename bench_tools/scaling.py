#!/usr/bin/env python
"""BASELINE.json configs 3 and 5 on N GPUs of one node (one process per GPU, torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench_tools/scaling.py [--frames 256] [--giga 32768] [--steps 5]

config 3  batch of 256 synthetic 1080p frames, RGB colour guide (3x3 covariance), r=16: frames are
          sharded over the ranks (contiguous blocks, NO collective), one gf_guided_batch launch per
          rank and step.  STRONG scaling: the batch is fixed, time = max over ranks.
config 5  one 32768x32768 gray image, r=16, sharded by row strips; per step every rank exchanges
          the 2r halo rows of I and p with its neighbours (NCCL send/recv over NVLink) and runs
          gf_guided_gray_strip on its strip.  Pixels come from a counter-based generator keyed by
          the GLOBAL (y, x), so any partition sees the same image; every rank re-computes a band
          around its upper seam from locally generated rows (no exchange) and compares.

Rank 0 prints one JSON line per config.  Developer/measurement tool: results are copied into
profiles/ by hand; bench.py remains the contract benchmark (config 2)."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402
from cudaimageprocessing_b200 import dist as D  # noqa: E402


def gen_rows(y0: int, y1: int, width: int, seed: int, device) -> torch.Tensor:
    """Rows [y0, y1) of the synthetic giga-image: a 32-bit integer hash of (y, x, seed) -> [0, 1)."""
    out = torch.empty((y1 - y0, width), device=device, dtype=torch.float32)
    x = torch.arange(width, device=device, dtype=torch.int64)
    for c0 in range(y0, y1, 1024):
        c1 = min(y1, c0 + 1024)
        y = torch.arange(c0, c1, device=device, dtype=torch.int64)[:, None]
        v = (y * 2654435761 + x * 40503 + seed * 97) & 0xFFFFFFFF
        v = ((v ^ (v >> 15)) * 2246822519) & 0xFFFFFFFF
        v = ((v ^ (v >> 13)) * 3266489917) & 0xFFFFFFFF
        v = v ^ (v >> 16)
        out[c0 - y0:c1 - y0] = (v >> 8).to(torch.float32) * (1.0 / 16777216.0)
    return out


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--giga", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--r", type=int, default=16)
    ap.add_argument("--skip3", action="store_true", help="skip the config-3 leg (diagnostics)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = pkg.api()
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    r, eps = args.r, 1e-2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- config 3: batch of 1080p colour-guide frames, sharded by frame ----------------
    if args.skip3:
        args.frames = world          # one frame per rank: the leg becomes negligible
    f0, f1 = D.shard_frames(args.frames, rank, world)
    n = f1 - f0
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    I = torch.rand((n, 1080, 1920, 3), device="cuda", generator=g)
    p = torch.rand((n, 1080, 1920), device="cuda", generator=g)
    q = torch.empty_like(p)
    for _ in range(2):
        D.filter_frames(api, I, p, q, r, eps, 0, sp)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        D.filter_frames(api, I, p, q, r, eps, 0, sp)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
    kern = api.last_kernel()
    if rank == 0:
        px = args.frames * 1080 * 1920
        print(json.dumps({"config": 3, "workload": f"{args.frames} x 1920x1080 RGB guide, 1-ch src, r={r}, eps=1e-2, REFLECT101",
                          "n_gpus": world, "frames_per_gpu": n, "ms_per_batch": ms, "mpix_s": px / ms / 1e3,
                          "alg_gb_s_per_gpu": 20.0 * px / world / ms / 1e6, "kernel": kern, "scaling": "strong", "collective": "none"}),
              flush=True)
    del I, p, q
    torch.cuda.empty_cache()

    # ---------------- config 5: one giga-image, row strips + 2r halo exchange ----------------
    Hh = Ww = args.giga
    y0, y1 = D.strip_rows(Hh, rank, world)
    bufI, viewI = D.alloc_strip(Hh, Ww, rank, world, r, "cuda")
    bufP, viewP = D.alloc_strip(Hh, Ww, rank, world, r, "cuda")
    viewI.copy_(gen_rows(y0, y1, Ww, 7, "cuda"))
    viewP.copy_(gen_rows(y0, y1, Ww, 8, "cuda"))
    qs = torch.empty((y1 - y0, Ww), device="cuda")
    ex0, ex1, k1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))

    def step():
        D.exchange_halos_inplace([bufI, bufP], Hh, rank, world, r)
        D.filter_strip(api, bufI, bufP, qs, Hh, rank, world, r, eps, 0, sp)
    for _ in range(2):
        step()
    barrier()
    t_ex = t_k = 0.0
    for _ in range(args.steps):
        barrier()
        ex0.record(stream)
        D.exchange_halos_inplace([bufI, bufP], Hh, rank, world, r)
        ex1.record(stream)
        D.filter_strip(api, bufI, bufP, qs, Hh, rank, world, r, eps, 0, sp)
        k1.record(stream)
        torch.cuda.synchronize()
        t_ex += ex0.elapsed_time(ex1)
        t_k += ex1.elapsed_time(k1)
    ms_ex = max_over_ranks(t_ex / args.steps, world)
    ms_k = max_over_ranks(t_k / args.steps, world)
    ms_tot = max_over_ranks((t_ex + t_k) / args.steps, world)
    kern = api.last_kernel()
    # the same step with the exchange hidden behind the interior rows (side stream)
    for _ in range(2):
        D.filter_strip_overlapped(api, bufI, bufP, qs, Hh, rank, world, r, eps, 0)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_ov = 0.0
    for _ in range(args.steps):
        barrier()
        o0.record(stream)
        D.filter_strip_overlapped(api, bufI, bufP, qs, Hh, rank, world, r, eps, 0)
        o1.record(stream)
        torch.cuda.synchronize()
        t_ov += o0.elapsed_time(o1)
    ms_ov = max_over_ranks(t_ov / args.steps, world)
    # seam check: a band around the upper seam, recomputed from locally generated rows without any exchange
    err = 0.0
    if rank > 0:
        b0, b1 = y0 - 64, y0 + 64
        lo, hi = b0 - 2 * r, b1 + 2 * r
        Il, Pl = gen_rows(lo, hi, Ww, 7, "cuda"), gen_rows(lo, hi, Ww, 8, "cuda")
        ql = torch.empty((b1 - b0, Ww), device="cuda")
        api.call("gf_guided_gray_strip", Il.data_ptr(), Pl.data_ptr(), ql.data_ptr(), Ww, Hh, lo, hi - lo, b0, b1 - b0, 0, 0, 0,
                 r, eps, 0, ctypes.c_void_p(sp))
        torch.cuda.synchronize()
        err = float((ql[64:] - qs[:64]).abs().max())      # rows [y0, y0+64) are this rank's
    errs = [err]
    if world > 1:
        t = torch.tensor([err], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        errs = [float(t.item())]
    if rank == 0:
        px = Hh * Ww
        print(json.dumps({"config": 5, "workload": f"{Ww}x{Hh} gray, r={r}, eps=1e-2, REFLECT101, row strips",
                          "n_gpus": world, "rows_per_gpu": y1 - y0, "ms_total": min(ms_ov, ms_tot), "ms_total_serial": ms_tot, "ms_total_overlapped": ms_ov,
                          "ms_halo_exchange": ms_ex, "ms_kernel": ms_k, "mpix_s": px / min(ms_ov, ms_tot) / 1e3, "alg_gb_s_per_gpu": 12.0 * px / world / ms_k / 1e6,
                          "halo_bytes_per_neighbour": 2 * r * Ww * 4 * 2, "seam_max_abs_diff_vs_local_recompute": errs[0],
                          "kernel": kern, "scaling": "strong", "collective": "NCCL send/recv (batch_isend_irecv); serial = exchange then one strip launch, overlapped = exchange on a side stream behind the interior rows + two seam launches"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
