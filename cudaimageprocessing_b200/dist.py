"""Multi-GPU partitioning of the guided filter (one process per GPU, torch.distributed).

The reference is single-GPU; the path shards in two ways:

* image batches (BASELINE config 3): frames are independent -> contiguous blocks of frames
  per rank, NO collective (`shard_frames`, `filter_frames`);
* one huge image (BASELINE config 5): row strips, rank g owns rows [g*H/G, (g+1)*H/G).  The
  fused single-pass kernel needs 2r rows of I and p beyond each seam, so neighbours exchange
  2r rows once per image with point-to-point send/recv (NCCL over NVLink on GPUs; gloo in the
  CPU tests) and every rank then runs `gf_guided_gray_strip` on its halo-extended buffer.
  The image's real top/bottom use the border rule inside the kernel.

Nothing here computes pixels: it only moves halos and calls the C ABI.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index: int):
    """Pins the calling process to the CPU cores next to GPU `device_index` (its NUMA node), so that pinned host buffers
    allocated afterwards (first touch) sit on the memory controller the GPU's PCIe root hangs off.  One process per GPU
    under torchrun otherwise runs wherever the scheduler puts it, and uploads from the far socket cross the CPU
    interconnect.  Returns the core list, or None when neither NVML nor sysfs knows it (nothing is changed then)."""
    import os
    cpus = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
    except Exception:
        cpus = None
    if not cpus:
        try:
            bdf = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
            if bdf:
                txt = open(f"/sys/bus/pci/devices/{bdf.lower()}/local_cpulist").read().strip()
                cpus = []
                for part in txt.split(","):
                    a, _, b = part.partition("-")
                    cpus += list(range(int(a), int(b or a) + 1))
        except Exception:
            cpus = None
    if not cpus:
        return None
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return allowed


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def strip_rows(height: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [y0, y1) of the image owned by `rank`."""
    return height * rank // world, height * (rank + 1) // world


def halo_rows(height: int, rank: int, world: int, r: int) -> Tuple[int, int]:
    """How many rows rank needs from the rank above / below (0 at the image's real border)."""
    y0, y1 = strip_rows(height, rank, world)
    top = 0 if rank == 0 else min(2 * r, y0)
    bot = 0 if rank == world - 1 else min(2 * r, height - y1)
    return top, bot


def alloc_strip(height: int, width: int, rank: int, world: int, r: int, device, dtype=torch.float32):
    """Buffer for a strip WITH room for its halos, and the view of the rows the rank owns.
    Filling the view and calling `exchange_halos_inplace` avoids any extra copy."""
    y0, y1 = strip_rows(height, rank, world)
    top, bot = halo_rows(height, rank, world, r)
    buf = torch.empty((top + (y1 - y0) + bot, width), device=device, dtype=dtype)
    return buf, buf[top:top + (y1 - y0)]


def exchange_halos_inplace(bufs, height: int, rank: int, world: int, r: int, group=None) -> int:
    """Fills the halo rows of every buffer in `bufs` (made by alloc_strip) from the neighbours.
    One batched round of isend/irecv (ncclSend/ncclRecv inside a group on GPUs).  Returns the
    global row index of row 0 of the buffers."""
    y0, y1 = strip_rows(height, rank, world)
    top, bot = halo_rows(height, rank, world, r)
    own = y1 - y0
    # validated from the geometry, identically on EVERY rank and before any P2P op exists: a rank that raised alone
    # would leave its neighbours waiting in batch_isend_irecv
    if world > 1:
        shortest = min(strip_rows(height, k, world)[1] - strip_rows(height, k, world)[0] for k in range(world))
        if shortest < 2 * r:
            raise ValueError(f"strips of {shortest} rows are shorter than the {2 * r}-row halo a neighbour needs")
    ops = []
    for b in bufs:
        if rank > 0:
            n_up = halo_rows(height, rank - 1, world, r)[1]        # rows the upper neighbour wants
            if n_up > own:
                raise ValueError(f"strip of {own} rows is shorter than the {n_up}-row halo its neighbour needs")
            ops.append(dist.P2POp(dist.isend, b[top:top + n_up], rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, b[0:top], rank - 1, group))
        if rank < world - 1:
            n_dn = halo_rows(height, rank + 1, world, r)[0]
            if n_dn > own:
                raise ValueError(f"strip of {own} rows is shorter than the {n_dn}-row halo its neighbour needs")
            ops.append(dist.P2POp(dist.isend, b[top + own - n_dn:top + own], rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, b[top + own:top + own + bot], rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return y0 - top


def exchange_halos(strip: torch.Tensor, height: int, rank: int, world: int, r: int, group=None):
    """Convenience form: takes the rank's own rows, returns (halo-extended buffer, buf_y0)."""
    buf, view = alloc_strip(height, strip.shape[1], rank, world, r, strip.device, strip.dtype)
    view.copy_(strip)
    buf_y0 = exchange_halos_inplace([buf], height, rank, world, r, group)
    return buf, buf_y0


def filter_strip(api, I_buf: torch.Tensor, p_buf: torch.Tensor, q_out: torch.Tensor, height: int, rank: int,
                 world: int, r: int, eps: float, border: int, stream: Optional[int] = None) -> None:
    """Runs the fused kernel on this rank's strip.  I_buf/p_buf: halo-extended buffers after the
    exchange (row 0 = global row y0 - top); q_out: the rank's own rows.  `stream` must be the stream the exchange was
    ordered on (torch's current stream for exchange_halos_inplace): the kernel reads the halo rows it wrote."""
    y0, y1 = strip_rows(height, rank, world)
    top, _ = halo_rows(height, rank, world, r)
    w = I_buf.shape[1]
    api.call("gf_guided_gray_strip", I_buf.data_ptr(), p_buf.data_ptr(), q_out.data_ptr(), w, height, y0 - top,
             I_buf.shape[0], y0, y1 - y0, I_buf.stride(0), p_buf.stride(0), q_out.stride(0), r, eps, border,
             ctypes.c_void_p(stream) if stream else None)


_SIDE_STREAMS = {}


def filter_strip_overlapped(api, I_buf: torch.Tensor, p_buf: torch.Tensor, q_out: torch.Tensor, height: int, rank: int,
                            world: int, r: int, eps: float, border: int, group=None) -> None:
    """Halo exchange + filter of this rank's strip with the exchange HIDDEN behind compute:
    the 2r halo rows travel on a side stream (NCCL send/recv over NVLink) while the current stream
    already filters the interior rows, whose 4r+1-row neighbourhood lies inside the rank's own rows;
    the two 2r-row seam bands are filtered once the halos have landed.  I_buf/p_buf come from
    alloc_strip with the rank's rows filled in; q_out holds the rank's own rows."""
    y0, y1 = strip_rows(height, rank, world)
    top, bot = halo_rows(height, rank, world, r)
    w = I_buf.shape[1]
    cuda = I_buf.is_cuda
    main = landed = None
    if cuda:
        main = torch.cuda.current_stream()
        dev = I_buf.device.index
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=I_buf.device)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            exchange_halos_inplace([I_buf, p_buf], height, rank, world, r, group)
            landed = torch.cuda.Event()
            landed.record(side)
    else:                              # host tensors (gloo tests): same row ranges, no overlap to be had
        exchange_halos_inplace([I_buf, p_buf], height, rank, world, r, group)

    def run(a, b):
        if b > a:
            api.call("gf_guided_gray_strip", I_buf.data_ptr(), p_buf.data_ptr(), q_out[a - y0:].data_ptr(), w, height, y0 - top,
                     I_buf.shape[0], a, b - a, I_buf.stride(0), p_buf.stride(0), q_out.stride(0), r, eps, border,
                     ctypes.c_void_p(main.cuda_stream) if cuda else None)
    lo = min(y1, y0 + 2 * r) if top else y0
    hi = max(lo, y1 - 2 * r) if bot else y1
    run(lo, hi)                       # interior: needs no halo row
    if cuda:
        main.wait_event(landed)
    run(y0, lo)                       # seam bands
    run(hi, y1)


# ---- row strips with the exchange behind the C ABI (gf_run_strips): peer-mapped strip buffers ----------------
class DeviceBuffer:
    """A gf_device_alloc'ed (cudaMalloc) float32 matrix: exportable over CUDA IPC, viewable as a torch tensor."""

    def __init__(self, api, rows: int, cols: int):
        self.api, self.rows, self.cols = api, rows, cols
        p = ctypes.c_void_p()
        api.call("gf_device_alloc", ctypes.addressof(p), rows * cols * 4)
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f4", "data": (self.ptr, False), "version": 2}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device="cuda")

    def ipc_handle(self) -> bytes:
        h = ctypes.create_string_buffer(64)
        self.api.call("gf_ipc_export", ctypes.c_void_p(self.ptr), h)
        return h.raw

    def free(self):
        if self.ptr:
            self.api.call("gf_device_free", ctypes.c_void_p(self.ptr))
            self.ptr = 0


class PeerStrips:
    """The strip buffers of one rank (guide + src, with room for the 2r halos) and the neighbours' buffers mapped
    into this process over CUDA IPC.  `run` is ONE C call per image: gf_run_strips pulls the halo rows from the
    neighbours with peer copies and launches the strip kernel (include/gf_b200.h)."""

    def __init__(self, api, height: int, width: int, rank: int, world: int, r: int, group=None):
        from ._capi import StripPeer
        self.api, self.height, self.width, self.rank, self.world, self.r = api, height, width, rank, world, r
        self.y0, self.y1 = strip_rows(height, rank, world)
        top, bot = ctypes.c_int(), ctypes.c_int()
        api.call("gf_strip_layout", height, self.y0, self.y1 - self.y0, r, ctypes.addressof(top), ctypes.addressof(bot))
        self.top, self.bot = top.value, bot.value
        rows = self.top + (self.y1 - self.y0) + self.bot
        self.guide, self.src = DeviceBuffer(api, rows, width), DeviceBuffer(api, rows, width)
        self.own_guide = self.guide.tensor()[self.top:self.top + (self.y1 - self.y0)]
        self.own_src = self.src.tensor()[self.top:self.top + (self.y1 - self.y0)]
        self._opened = []
        self.up = self.down = None
        if world > 1:
            mine = (self.guide.ipc_handle(), self.src.ipc_handle(), self.top, self.y1 - self.y0)
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)

            def peer(k):
                hg, hs, ptop, prows = everyone[k]
                pg, ps = ctypes.c_void_p(), ctypes.c_void_p()
                api.call("gf_ipc_open", hg, ctypes.addressof(pg))
                api.call("gf_ipc_open", hs, ctypes.addressof(ps))
                self._opened += [pg.value, ps.value]
                return StripPeer(pg.value, ps.value, width, width, ptop, prows)
            if rank > 0:
                self.up = peer(rank - 1)
            if rank < world - 1:
                self.down = peer(rank + 1)

    def run(self, q_out: torch.Tensor, eps: float, border: int, stream: Optional[int] = None) -> None:
        self.api.call("gf_run_strips", ctypes.c_void_p(self.guide.ptr), ctypes.c_void_p(self.src.ptr), q_out.data_ptr(), self.width,
                      self.height, self.y0, self.y1 - self.y0, self.width, self.width, q_out.stride(0), self.r, eps, border,
                      ctypes.byref(self.up) if self.up is not None else None,
                      ctypes.byref(self.down) if self.down is not None else None,
                      ctypes.c_void_p(stream) if stream else None)

    def close(self):
        for p in self._opened:
            self.api.call("gf_ipc_close", ctypes.c_void_p(p))
        self._opened = []
        self.guide.free()
        self.src.free()


def filter_frames(api, I: torch.Tensor, p: torch.Tensor, q: torch.Tensor, r: int, eps: float, border: int,
                  stream: Optional[int] = None) -> None:
    """Filters this rank's block of frames with ONE launch.  I: [n,h,w] (gray) or [n,h,w,3]
    (colour guide); p, q: [n,h,w]."""
    n, h, w = p.shape
    gch = 1 if I.dim() == 3 else I.shape[3]
    api.call("gf_guided_batch", I.data_ptr(), p.data_ptr(), q.data_ptr(), n, w, h, gch, 0, 0, 0, 0, 0, 0, r, eps,
             border, ctypes.c_void_p(stream) if stream else None)
