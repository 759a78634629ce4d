"""Developer tool (torchrun, >= 2 GPUs): time of the 2r-row halo exchange of a 32768-wide strip (NCCL send/recv)."""
import os, sys, json, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cudaimageprocessing_b200 import dist as D
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, r = 4096 * world, 32768, 16
bufI, _ = D.alloc_strip(H, W, rank, world, r, "cuda"); bufP, _ = D.alloc_strip(H, W, rank, world, r, "cuda")
bufI.zero_(); bufP.zero_()
for _ in range(5): D.exchange_halos_inplace([bufI, bufP], H, rank, world, r)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): D.exchange_halos_inplace([bufI, bufP], H, rank, world, r)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "ms_exchange": float(t.item()), "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}), flush=True)
dist.destroy_process_group()
