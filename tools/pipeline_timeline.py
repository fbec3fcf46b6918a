#!/usr/bin/env python
"""Timeline of the multi-GPU sort pipeline (development aid): runs a few jobs through DistributedSortPipeline with
GLU_PIPE_TRACE=1 and prints, for rank 0, when every phase of every job ended on the device (ms after the first traced
job's start) next to the host's own clock.

  torchrun --nproc-per-node N tools/pipeline_timeline.py [--log2-pairs 28] [--jobs 8]
"""
import argparse
import os
import sys
import time

os.environ["GLU_PIPE_TRACE"] = "1"
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--log2-pairs", type=int, default=28)
p.add_argument("--jobs", type=int, default=8)
args = p.parse_args()
glu = entry.load_package()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << args.log2_pairs
g = torch.Generator(device=dev).manual_seed(1 + rank)
inputs = [(torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g),
           torch.arange(n, dtype=torch.int32, device=dev)) for _ in range(2)]
pipe = glu.DistributedSortPipeline(n)
for i in range(4):  # warm-up
    pipe.submit(*inputs[i % 2], n)
pipe.flush()
torch.cuda.synchronize()
dist.barrier()
pipe.trace.clear()
t0 = torch.cuda.Event(enable_timing=True)
t0.record()
h0 = time.perf_counter()
for i in range(args.jobs):
    pipe.submit(*inputs[i % 2], n)
pipe.flush()
t1 = torch.cuda.Event(enable_timing=True)
t1.record()
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    lane = pipe.lanes[0]
    print(f"world {world}, 2^{args.log2_pairs} pairs per GPU, {pipe.num_lanes} lanes, style {lane.exchange_style}, "
          f"sync {lane._dma_sync}, copy stream {'yes' if pipe.stream_d is not None else 'no'}: "
          f"{t0.elapsed_time(t1) / args.jobs:.3f} ms per job")
    cols = ["x_begin", "hist", "plan", "msd", "x_end", "copies_end", "s_begin", "s_end"]
    print("job | host: submit  hist_seen  plan_done  return | device: " + "  ".join(f"{c:>10s}" for c in cols))
    for tr in pipe.trace:
        host = [1e3 * (tr.get(k, float("nan")) - h0) for k in ("host_submit", "host_hist_seen", "host_plan_done", "host_return")]
        devt = [t0.elapsed_time(tr[c]) if c in tr else float("nan") for c in cols]
        print(f"{tr['job'] - pipe.trace[0]['job']:3d} | " + "  ".join(f"{h:9.3f}" for h in host) + " | " +
              "  ".join(f"{d:10.3f}" for d in devt))
pipe.close()
dist.destroy_process_group()
