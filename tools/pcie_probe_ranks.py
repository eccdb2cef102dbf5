#!/usr/bin/env python
"""Host-link ceiling of a box with N ranks copying at once (the limit of bench.py's `e2e` figure at N GPUs):
    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe_ranks.py
Every rank copies a 256-tile batch (268 MB) from and to page-locked host memory on two streams, all ranks together
between barriers; rank 0 prints per-rank and aggregate GB/s per direction."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 * 512 * 512
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_a, d_b = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(2):
    both()
barrier()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    both()
barrier()
dt = (time.perf_counter() - t0) / reps
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    dt = float(t[0])
    print(json.dumps({"ranks": world, "bytes_per_rank_per_direction": n * 4, "ms": 1e3 * dt,
                      "GBps_per_rank_per_direction": gb / dt, "GBps_aggregate_per_direction": world * gb / dt,
                      "implied_e2e_tiles_per_s_fp32_io": world * 256 / dt}))
if world > 1:
    dist.destroy_process_group()
