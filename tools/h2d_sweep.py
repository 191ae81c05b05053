#!/usr/bin/env python3
"""Under torchrun, ON THE GPU BOX: what the host can feed N GPUs at once.  Every rank copies a pinned 1 GiB buffer to its
GPU (and, second pass, back; third pass, both directions at once on two streams), all ranks together between barriers;
rank 0 prints the aggregate GB/s.  This is the ceiling of the end-to-end scaling curve.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_sweep.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(kind, reps=6):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_dir = n * reps * world / t.item() / 1e9
    return per_dir * (2 if kind == "both" else 1)


run("h2d", 2)
out = {"kind": "pcie_sweep", "n_gpus": world, "bytes_per_copy": n, "h2d_GBps": run("h2d"), "d2h_GBps": run("d2h"), "both_GBps_total": run("both")}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
