#!/usr/bin/env python
"""Measured NVLink peer-store peak (run under torch.distributed.run): every rank fills 1 GiB of rank (r+1)%P's memory
through its CUDA-IPC mapping with SM-issued, fully coalesced, line-aligned 4-byte stores (glu_memset_u32's fill kernel)
— the best case of what the exchange pass of the multi-GPU sort does — all ranks at once, so every GPU sends and
receives at the same time.  Also the same fill on local memory (HBM write peak) and a DMA peer copy (cudaMemcpyPeer).
Prints one JSON line on rank 0: the denominators for `partition_exchange` fractions in profiles/."""
import ctypes
import json
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
n = 1 << 28  # 1 GiB of uint32

# a cudaMalloc'ed (IPC-exportable) buffer per rank, mapped by everybody
buf = ctypes.c_void_p()
glu.check(glu.lib.glu_malloc(ctypes.byref(buf), 4 * n), "glu_malloc")
h = ctypes.create_string_buffer(64)
glu.check(glu.lib.glu_ipc_get_handle(buf, h), "glu_ipc_get_handle")
handles = [None] * world
dist.all_gather_object(handles, h.raw)
nxt = (rank + 1) % world
peer = ctypes.c_void_p()
if world > 1:
    glu.check(glu.lib.glu_ipc_open_handle(handles[nxt], ctypes.byref(peer)), "glu_ipc_open_handle")
else:
    peer = buf
st = int(torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


local_ms = timed(lambda: glu.check(glu.lib.glu_memset_u32(buf, 0x01020304, n, st), "fill local"))
peer_ms = timed(lambda: glu.check(glu.lib.glu_memset_u32(peer, 0x01020304, n, st), "fill peer"))
src = torch.empty(n, dtype=torch.int32, device=dev)
dma_ms = timed(lambda: glu.check(glu.lib.glu_memcpy_d2d(peer, src.data_ptr(), 4 * n, st), "memcpy peer"))
if rank == 0:
    print(json.dumps({"world": world, "bytes": 4 * n,
                      "hbm_fill_GB/s": 4 * n / local_ms / 1e6,
                      "nvlink_peer_store_GB/s_per_gpu_per_direction": 4 * n / peer_ms / 1e6,
                      "nvlink_peer_dma_GB/s_per_gpu_per_direction": 4 * n / dma_ms / 1e6,
                      "how": "every rank writes 1 GiB into rank (r+1)%P at the same time; max over ranks"}), flush=True)
dist.barrier()
if world > 1:
    glu.lib.glu_ipc_close_handle(peer)
dist.barrier()
glu.lib.glu_free(buf)
dist.destroy_process_group()
