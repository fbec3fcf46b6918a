#!/usr/bin/env python
"""Per-phase timing of DistributedRadixSort (run under torch.distributed.run); development aid."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

glu = entry.load_package()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << int(os.environ.get("LOG2N", "28"))
g = torch.Generator(device=dev).manual_seed(1 + rank)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
vals = torch.arange(n, dtype=torch.int32, device=dev)
for exchange in os.environ.get("EXCHANGES", "p2p,nccl").split(","):
    sorter = glu.DistributedRadixSort(n, exchange=exchange)
    for _ in range(3):
        sorter(keys, vals, n)
    torch.cuda.synchronize()
    dist.barrier()
    reps = 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        sorter(keys, vals, n)
    b.record()
    torch.cuda.synchronize()
    total = a.elapsed_time(b) / reps
    sorter.timing = {}
    for _ in range(reps):
        sorter(keys, vals, n)
    phases = {k: v / reps for k, v in sorter.timing.items()}
    sorter.timing = None
    if rank == 0:
        print(f"world {world} {exchange}: {total:.3f} ms/step unsynchronised; phases (synchronised, ms): "
              + ", ".join(f"{k} {v:.3f}" for k, v in phases.items()), flush=True)
    dist.barrier()
dist.destroy_process_group()
