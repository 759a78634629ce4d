"""CPU tests of the N>1 host logic with torch.distributed / gloo, world_size 2 and 3:
frame sharding, strip partitioning, and the halo exchange (the same isend/irecv code that runs
over NCCL on GPUs).  The exchanged buffers are then filtered strip by strip -- through the
test-only emulator build of the kernels, since there is no GPU here -- and the stitched result
must equal the oracle on the whole image."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, r, border, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cudaimageprocessing_b200 import dist as D
        rng = np.random.default_rng(77)
        I = rng.random((H, W), dtype=np.float32)
        p = rng.random((H, W), dtype=np.float32)
        y0, y1 = D.strip_rows(H, rank, world)
        bufI, viewI = D.alloc_strip(H, W, rank, world, r, "cpu")
        bufP, viewP = D.alloc_strip(H, W, rank, world, r, "cpu")
        bufI.fill_(float("nan")); bufP.fill_(float("nan"))
        viewI.copy_(torch.from_numpy(I[y0:y1])); viewP.copy_(torch.from_numpy(p[y0:y1]))
        buf_y0 = D.exchange_halos_inplace([bufI, bufP], H, rank, world, r)
        top, bot = D.halo_rows(H, rank, world, r)
        assert buf_y0 == y0 - top
        assert np.array_equal(bufI.numpy(), I[y0 - top:y1 + bot]), "halo exchange delivered wrong rows"
        assert np.array_equal(bufP.numpy(), p[y0 - top:y1 + bot])
        b2, y2 = D.exchange_halos(torch.from_numpy(I[y0:y1].copy()), H, rank, world, r)
        assert y2 == buf_y0 and np.array_equal(b2.numpy(), bufI.numpy())
        # filter the strip with the kernels under the emulator (test infrastructure)
        from gf_backend import EmuBackend
        be = EmuBackend()
        q = torch.empty((y1 - y0, W), dtype=torch.float32)
        D.filter_strip(be.api, bufI, bufP, q, H, rank, world, r, 1e-2, border)
        np.save(os.path.join(out_dir, f"q_{rank}.npy"), q.numpy())
        # the overlapped form (interior rows first, seam bands after the halos) gives the same strip
        bufI2, viewI2 = D.alloc_strip(H, W, rank, world, r, "cpu")
        bufP2, viewP2 = D.alloc_strip(H, W, rank, world, r, "cpu")
        viewI2.copy_(torch.from_numpy(I[y0:y1])); viewP2.copy_(torch.from_numpy(p[y0:y1]))
        q2 = torch.full((y1 - y0, W), float("nan"), dtype=torch.float32)
        D.filter_strip_overlapped(be.api, bufI2, bufP2, q2, H, rank, world, r, 1e-2, border)
        assert np.abs(q2.numpy() - q.numpy()).max() <= 2e-6
        # batch sharding covers every frame exactly once
        spans = [D.shard_frames(11, k, world) for k in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 11 and all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,r,border", [(2, 40, 36, 3, 0), (3, 45, 52, 2, 1), (2, 24, 40, 4, 2)])
def test_strips_over_gloo(tmp_path, world, H, W, r, border):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, W, r, border, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import gf_oracle as O
    rng = np.random.default_rng(77)
    I = rng.random((H, W), dtype=np.float32)
    p = rng.random((H, W), dtype=np.float32)
    q = np.concatenate([np.load(tmp_path / f"q_{k}.npy") for k in range(world)], axis=0)
    assert np.abs(q - O.guided_filter_gray(I, p, r, 1e-2, border, np.float64)).max() <= 1e-4


def test_strip_too_short_is_refused():
    from cudaimageprocessing_b200 import dist as D
    assert D.halo_rows(100, 0, 4, 8) == (0, 16) and D.halo_rows(100, 3, 4, 8) == (16, 0)
    assert D.strip_rows(10, 1, 3) == (3, 6)


def test_bind_host_to_gpu_without_a_gpu_changes_nothing():
    """bind_host_to_gpu (bench.py, N > 1) must leave the process alone when neither NVML nor sysfs knows the GPU"""
    import os
    from cudaimageprocessing_b200 import dist as D
    before = os.sched_getaffinity(0)
    cpus = D.bind_host_to_gpu(0)
    after = os.sched_getaffinity(0)
    assert cpus is None and after == before or (cpus is not None and set(cpus) == after and after <= before)
