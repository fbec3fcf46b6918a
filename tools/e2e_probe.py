#!/usr/bin/env python
"""What bounds bench.py's e2e number: PCIe copy rates of this box (H2D alone, D2H alone, both at once) next to the
host sort queue's pairs/s at several depths (development aid; bench.py is the contract)."""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--log2n", type=int, default=28)
p.add_argument("--jobs", type=int, default=6)
p.add_argument("--depths", default="1,2,3,4")
args = p.parse_args()
glu = entry.load_package()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n = 1 << args.log2n
nbytes = 4 * n

h_in = torch.empty(2 * n, dtype=torch.int32).pin_memory()
h_out = torch.empty(2 * n, dtype=torch.int32).pin_memory()
d_a = torch.empty(2 * n, dtype=torch.int32, device=dev)
d_b = torch.empty(2 * n, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def wall(fn, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


gb = 2 * nbytes / 1e9
t_h2d, t_d2h, t_both = wall(h2d), wall(d2h), wall(both)
print(f"pcie {gb:.2f} GB per direction: H2D alone {gb / t_h2d:.1f} GB/s ({1e3 * t_h2d:.1f} ms), D2H alone {gb / t_d2h:.1f} GB/s "
      f"({1e3 * t_d2h:.1f} ms), both at once {gb / t_both:.1f} GB/s per direction ({1e3 * t_both:.1f} ms)")
print(f"     => e2e ceiling of a 2^{args.log2n}-pair sort step: {n / t_both / 1e9:.2f} Gpairs/s (copies fully overlapped), "
      f"{n / (t_h2d + t_d2h) / 1e9:.2f} Gpairs/s (one call, nothing overlapped, sort time excluded)")
del d_a, d_b, h_in, h_out


def pinned_u32(count):
    ptr = ctypes.c_void_p()
    glu.check(glu.lib.glu_malloc_host(ctypes.byref(ptr), 4 * count), "glu_malloc_host")
    buf = (ctypes.c_uint32 * count).from_address(ptr.value)
    return np.frombuffer(buf, dtype=np.uint32), ptr


jobs = []
for i in range(args.jobs):
    hk, pk = pinned_u32(n)
    hv, pv = pinned_u32(n)
    jobs.append((hk, hv, pk, pv))
for depth in [int(x) for x in args.depths.split(",")]:
    for i, (hk, hv, _, _) in enumerate(jobs):
        hk[:] = np.random.default_rng(100 * depth + i).integers(0, 1 << 32, size=n, dtype=np.uint32)
        hv[:] = np.arange(n, dtype=np.uint32)
    q = glu.HostSortQueue(n, depth=depth)
    q.submit(jobs[0][0][: 1 << 20].copy(), jobs[0][1][: 1 << 20].copy())  # warm the streams (pageable, tiny)
    q.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for hk, hv, _, _ in jobs:
        q.submit(hk, hv, n)
    q.wait()
    t = time.perf_counter() - t0
    q.close()
    ok = all(bool(np.all(hk[:-1][: 1 << 22] <= hk[1:][: 1 << 22])) for hk, _, _, _ in jobs)
    print(f"host sort queue depth {depth}: {args.jobs} jobs of 2^{args.log2n} pairs in {1e3 * t:.1f} ms = {1e3 * t / args.jobs:.1f} ms/job, "
          f"{n * args.jobs / t / 1e9:.2f} Gpairs/s  sorted={ok}")
