#!/usr/bin/env python
"""Aggregate PCIe ceiling of the box with ALL ranks copying at once (run under torch.distributed.run): every rank moves
2 GiB host->device and 2 GiB device->host from / to pinned memory, alone and both directions together; barriers around
every measurement, max over ranks.  What bounds bench.py's e2e number at N GPUs (each step moves 8 B per pair each way).
Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << 29  # 2 GiB of int32
h_in = torch.empty(n, dtype=torch.int32).pin_memory()
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
h_in.zero_()
d_a = torch.empty(n, dtype=torch.int32, device=dev)
d_b = torch.zeros(n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def wall(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


gb = 4 * n / 1e9
t1, t2, t3 = wall(h2d), wall(d2h), wall(both)
if rank == 0:
    print(json.dumps({"world": world, "GB_per_rank_per_direction": gb,
                      "h2d_alone_GB/s_aggregate": world * gb / t1, "d2h_alone_GB/s_aggregate": world * gb / t2,
                      "both_GB/s_aggregate_per_direction": world * gb / t3,
                      "e2e_ceiling_Gpairs/s": world * (n / 2) / t3 / 1e9,
                      "how": "all ranks copy 2 GiB each way at once from / to pinned host memory; wall clock between "
                             "barriers, max over ranks; ceiling = pairs whose 8 B go each way in that time"}), flush=True)
dist.destroy_process_group()
