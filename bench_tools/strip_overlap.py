"""Developer tool (torchrun, N GPUs): BASELINE configs[4] (32768^2 gray, r=16, row strips) with the halo exchange behind
the strip kernel (gf_run_strips' default) and with the sequential form (pull, then one launch) beside it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        bench_tools/strip_overlap.py > gpurun_out/strip_overlap.jsonl

One JSON line per setting (rank 0): ms_total = barrier -> pull + kernel(s) done, max over ranks, CUDA events."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudaimageprocessing_b200 as pkg  # noqa: E402
from bench_tools import legs  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api = pkg.api()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    size = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    settings = [("pull_then_launch", {"GF_STRIP_OVERLAP": 0}), ("seams_behind_main", {"GF_STRIP_OVERLAP": 1}),
                ("pull_then_launch", {"GF_STRIP_OVERLAP": 0}), ("seams_behind_main", {"GF_STRIP_OVERLAP": 1})]
    for name, opts in settings:
        for k, v in opts.items():
            api.set_option(k, v)
        out = legs.config5(torch, dist, api, pkg, rank, world, barrier, size=size, steps=5)
        if rank == 0:
            print(json.dumps({"setting": name, "n_gpus": world, "ms_total": round(out["ms_total"], 4), "ms_kernel_alone": round(out["ms_kernel"], 4),
                              "seam_err": out["seam_check"]["max_abs_err_vs_oracle_f64"], "kernel": out["kernel"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
