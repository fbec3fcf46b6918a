#!/usr/bin/env python
"""Parity check + timing of DistributedSortPipeline (exchange of job k+1 under the local sort of job k) against
back-to-back DistributedRadixSort calls (development aid; the same checks run in tests/test_multigpu_gpu.py and the
numbers come from bench.py).

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29733 tools/pipeline_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402
import oracle  # noqa: E402  (checker only)

glu = entry.load_package()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()


def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


# ---- parity: four small jobs in flight two at a time, each checked against std::stable_sort of the global input
n = 300_007 + 1013 * rank
pipe = glu.DistributedSortPipeline(400_000, capacity_factor=2.5)
jobs, tickets = [], []
for j in range(4):
    keys = oracle.mt19937_u32(100 * j + rank, n)
    base = sum(gather(n)[:rank])
    vals = np.arange(base, base + n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
    dv = torch.from_numpy(vals.view(np.int32).copy()).to(dev)
    jobs.append((keys, vals, dk, dv))
    tickets.append(pipe.submit(dk, dv, n))
    if j >= 1:  # the previous job's result must be taken before the next submit reuses its lane
        sk, sv, m = pipe.result(tickets[j - 1])
        torch.cuda.synchronize()
        res = gather((jobs[j - 1][0], jobs[j - 1][1], sk.cpu().numpy().view(np.uint32), sv.cpu().numpy().view(np.uint32)))
        if rank == 0:
            ek, ev = oracle.stable_sort_pairs(np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]))
            assert np.array_equal(np.concatenate([r[2] for r in res]), ek), f"job {j - 1}: keys differ"
            assert np.array_equal(np.concatenate([r[3] for r in res]), ev), f"job {j - 1}: values differ"
if rank == 0:
    print("pipeline parity OK (3 jobs checked)", flush=True)
del pipe, jobs

# ---- timing at 2^28 pairs per GPU: back-to-back sorts vs the pipeline
n = 1 << int(os.environ.get("LOG2N", "28"))
g = torch.Generator(device=dev).manual_seed(1 + rank)
inputs = [(torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g),
           torch.arange(n, dtype=torch.int32, device=dev)) for _ in range(4)]
reps = 10
plain = glu.DistributedRadixSort(n)
for k, v in inputs[:3]:
    plain(k, v, n)
torch.cuda.synchronize()
dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(reps):
    plain(*inputs[i % 4], n)
b.record()
torch.cuda.synchronize()
t_plain = a.elapsed_time(b) / reps
del plain
torch.cuda.empty_cache()
pipe = glu.DistributedSortPipeline(n)
for k, v in inputs[:3]:
    pipe.submit(k, v, n)
pipe.flush()
torch.cuda.synchronize()
dist.barrier()
a.record()
for i in range(reps):
    pipe.submit(*inputs[i % 4], n)
pipe.flush()
b.record()
torch.cuda.synchronize()
t_pipe = a.elapsed_time(b) / reps
sk, _, m = pipe.result(pipe._submitted - 1)
k64 = sk[: 1 << 22].to(torch.int64) & 0xFFFFFFFF
assert bool((k64[1:] >= k64[:-1]).all()), "pipeline output is not sorted"
if rank == 0:
    print(f"world {world}, 2^{n.bit_length() - 1} pairs per GPU: back to back {t_plain:.3f} ms/job "
          f"({world * n / t_plain / 1e6:.1f} Gpairs/s), pipelined {t_pipe:.3f} ms/job ({world * n / t_pipe / 1e6:.1f} Gpairs/s)",
          flush=True)
dist.destroy_process_group()
