#!/usr/bin/env python
"""tools/h2d_ceiling.py — what the BOX gives for host -> device copies, without a line of this library: every rank copies
a pinned 2 GiB buffer to its GPU with plain cudaMemcpyAsync (torch), all ranks at once, barrier-bracketed.  Run under
torchrun with 1, 2, 4, 8 ranks; the aggregate is the ceiling `e2e` (which re-ingests all observations every step) can reach.
  torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 28                                  # 2 GiB of fp64
host = torch.empty(n, dtype=torch.float64).pin_memory()
host.fill_(1.0)
dev = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 8
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    gb = n * 8 * reps / 1e9
    print(json.dumps({"ranks": world, "h2d_gbs_per_gpu": gb / float(dt), "h2d_gbs_aggregate": gb * world / float(dt),
                      "what": "pinned host -> device, 2 GiB x 8 per rank, plain cudaMemcpyAsync via torch, all ranks concurrently"}), flush=True)
if world > 1:
    dist.destroy_process_group()
