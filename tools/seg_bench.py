#!/usr/bin/env python
"""Device-side timing of the segmented sort (the local step of the multi-GPU sort) on one GPU: 32 buckets of 2^23 pairs,
key bits [0, 24) — the histogram kernel and the digit passes separately (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

glu = entry.load_package()
dev = torch.device("cuda", 0)
tile = int(glu.lib.glu_radix_sort_segment_tile())
segs, per = 32, 1 << 23
counts = np.full(segs, per, dtype=np.uint32)
tiles = -(-per // tile)
max_tiles = segs * tiles + 8
g = torch.Generator(device=dev).manual_seed(1)
ka = torch.randint(-(1 << 31), (1 << 31) - 1, (max_tiles * tile,), dtype=torch.int32, device=dev, generator=g)
va = torch.arange(max_tiles * tile, dtype=torch.int32, device=dev)
kb, vb = torch.empty_like(ka), torch.empty_like(va)
dc = torch.from_numpy(counts.view(np.int32)).to(dev)
sorter = glu.RadixSort()
glu.profile_enable(True)
for i in range(6):
    if i == 2:
        torch.cuda.synchronize()
        glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
        glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    sorter.sort_segmented(ka, va, kb, vb, dc, segs, max_tiles, 0, 24)
b.record()
torch.cuda.synchronize()
h_ms, h_n = glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
s_ms, s_n = glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
n = segs * per
print(f"segmented sort, {segs} x 2^23 pairs, bits [0,24): {a.elapsed_time(b) / 4:.3f} ms per sort; histogram {h_ms / h_n:.3f} ms "
      f"({4 * n / (h_ms / h_n) / 1e6:.0f} GB/s), pass {s_ms / s_n:.3f} ms x 3")
