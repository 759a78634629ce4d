"""Developer tool (torchrun, N GPUs): per-rank timings of the config-5 strip step, with and without the
halo exchange in front of the kernel, to localise a slow rank / slow phase."""
import ctypes, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg
from cudaimageprocessing_b200 import dist as D
from bench_tools.scaling import gen_rows

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
api = pkg.api()
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
H = W = 32768; r = 16
y0, y1 = D.strip_rows(H, rank, world)
bufI, viewI = D.alloc_strip(H, W, rank, world, r, "cuda")
bufP, viewP = D.alloc_strip(H, W, rank, world, r, "cuda")
qs = torch.empty((y1 - y0, W), device="cuda")
res = {"rank": rank, "rows": y1 - y0, "buf_rows": bufI.shape[0], "ptr_mod_2M": [bufI.data_ptr() % (1 << 21), bufP.data_ptr() % (1 << 21), qs.data_ptr() % (1 << 21)]}


def t_kernel(n=5):
    D.filter_strip(api, bufI, bufP, qs, H, rank, world, r, 1e-2, 0, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        D.filter_strip(api, bufI, bufP, qs, H, rank, world, r, 1e-2, 0, sp)
    e1.record(stream); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)


bufI.uniform_(); bufP.uniform_()
res["kernel_rand_ms"] = t_kernel()
bufI.zero_(); bufP.zero_()
viewI.copy_(gen_rows(y0, y1, W, 7, "cuda")); viewP.copy_(gen_rows(y0, y1, W, 8, "cuda"))
res["kernel_hash_halo0_ms"] = t_kernel()
dist.barrier(); torch.cuda.synchronize()
D.exchange_halos_inplace([bufI, bufP], H, rank, world, r)
torch.cuda.synchronize()
res["kernel_hash_after_exchange_ms"] = t_kernel()


def loop(sync_between, n=5):
    ex0, ex1, k1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_ex = t_k = 0.0
    for _ in range(n):
        dist.barrier(); torch.cuda.synchronize()
        ex0.record(stream)
        D.exchange_halos_inplace([bufI, bufP], H, rank, world, r)
        ex1.record(stream)
        if sync_between:
            torch.cuda.synchronize()
            ex1.record(stream)
        D.filter_strip(api, bufI, bufP, qs, H, rank, world, r, 1e-2, 0, sp)
        k1.record(stream)
        torch.cuda.synchronize()
        t_k += ex1.elapsed_time(k1)
    return round(t_k / n, 3)


res["loop_kernel_ms"] = loop(False)
res["loop_kernel_sync_between_ms"] = loop(True)
res["loop_kernel_again_ms"] = loop(False)
# the allocation pattern of scaling.py: a large batch allocated and released before the strip buffers exist
big = torch.rand((64, 1080, 1920, 3), device="cuda"); big2 = torch.rand((64, 1080, 1920), device="cuda")
del big, big2
torch.cuda.empty_cache()
res["loop_kernel_after_alloc_free_ms"] = loop(False)
res["kernel_alone_end_ms"] = t_kernel()
res["nan_in_bufs"] = int(torch.isnan(bufI).sum() + torch.isnan(bufP).sum())
res["absmax"] = float(max(bufI.abs().max(), bufP.abs().max()))
res["kernel"] = api.last_kernel()
out = [None] * world
dist.all_gather_object(out, res)
if rank == 0:
    for o in out:
        print(json.dumps(o), flush=True)
dist.destroy_process_group()
