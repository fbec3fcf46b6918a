#!/usr/bin/env python
"""Quick device-side timing of the three primitives (development aid; bench.py is the contract)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--log2n", type=int, default=28)
p.add_argument("--reps", type=int, default=10)
p.add_argument("--what", default="sort,scan,reduce")
p.add_argument("--dist", default="uniform")
args = p.parse_args()

glu = entry.load_package()
dev = torch.device("cuda", 0)
n = 1 << args.log2n
g = torch.Generator(device=dev).manual_seed(1)
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, prep=None, reps=args.reps):
    times = []
    for i in range(reps + 3):
        if prep:
            prep()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2], times[0]


if "sort" in args.what:
    if args.dist == "uniform":
        keys0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    elif args.dist == "zero":
        keys0 = torch.zeros(n, dtype=torch.int32, device=dev)
    elif args.dist == "ent16":
        keys0 = torch.randint(0, 1 << 16, (n,), dtype=torch.int32, device=dev, generator=g)
    elif args.dist == "ent16hi":
        keys0 = torch.randint(0, 1 << 16, (n,), dtype=torch.int32, device=dev, generator=g) << 16
    elif args.dist == "zipf":
        # Zipf(s = 1.1) over 2^20 distinct keys by inverse-CDF sampling (BASELINE.json configs[4])
        ranks = torch.arange(1, (1 << 20) + 1, dtype=torch.float64, device=dev)
        cdf = torch.cumsum(ranks.pow(-1.1), 0)
        cdf /= cdf[-1].clone()
        u = torch.rand(n, dtype=torch.float64, device=dev, generator=g)
        keys0 = torch.searchsorted(cdf, u).to(torch.int32)
        del ranks, cdf, u
    else:
        raise SystemExit(f"unknown --dist {args.dist}")
    vals0 = torch.arange(n, dtype=torch.int32, device=dev)
    keys, vals = keys0.clone(), vals0.clone()
    sorter = glu.RadixSort()
    sorter.prepare_internal_buffers(n)

    def prep():
        keys.copy_(keys0)
        vals.copy_(vals0)

    med, best = timeit(lambda: sorter(keys, vals, n), prep)
    glu.profile_enable(True)
    prep()
    sorter(keys, vals, n)
    torch.cuda.synchronize()
    sweep_ms, sweep_n = glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
    hist_ms, hist_n = glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
    glu.profile_enable(False)
    print(f"       histogram {hist_ms / max(1, hist_n):.3f} ms, onesweep pass {sweep_ms / max(1, sweep_n):.3f} ms x {sweep_n}")
    print(f"sort   n=2^{args.log2n} {args.dist}: median {med:.3f} ms  best {best:.3f} ms  "
          f"{n / med / 1e6:.2f} Gpairs/s  {68 * n / med / 1e6:.0f} GB/s(68B/pair)  "
          f"cfg={os.environ.get('GLU_SORT_CONFIG', 'auto')} rank={os.environ.get('GLU_SORT_RANK', '0')} "
          f"tma={os.environ.get('GLU_SORT_TMA', '1')} prefetch={os.environ.get('GLU_SORT_PREFETCH', 'auto')} "
          f"options={os.environ.get('GLU_SORT_OPTIONS', '0')}")
    if not int(os.environ.get("GLU_SORT_OPTIONS", "0")) & 1:  # bit 0 = look-back skipped (timing experiment, wrong results)
        k64 = keys.to(torch.int64) & 0xFFFFFFFF
        assert bool((k64[1:] >= k64[:-1]).all()), "not sorted"
        del k64

if "ex" in args.what.split(","):
    # glu_radix_sort_u32_ex flavours (uniform keys): key-only (36 B/key), descending pairs, low 24 bits only
    keys0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    vals0 = torch.arange(n, dtype=torch.int32, device=dev)
    keys, vals = keys0.clone(), vals0.clone()
    sorter = glu.RadixSort()

    def prep_ex():
        keys.copy_(keys0)
        vals.copy_(vals0)

    for name, fn, bytes_per in [
            ("keys only", lambda: sorter.sort_ex(keys, None, n), 36),
            ("keys only descending", lambda: sorter.sort_ex(keys, None, n, 0, 32, True), 36),
            ("pairs descending", lambda: sorter.sort_ex(keys, vals, n, 0, 32, True), 68),
            ("pairs bits [0,24)", lambda: sorter.sort_ex(keys, vals, n, 0, 24), 52 + 16),  # + the copy back (odd passes)
            ("pairs bits [8,32)", lambda: sorter.sort_ex(keys, vals, n, 8, 32), 52 + 16)]:
        med, best = timeit(fn, prep_ex)
        print(f"sort_ex {name:22s} n=2^{args.log2n}: median {med:.3f} ms  best {best:.3f} ms  {n / med / 1e6:.2f} Gkeys/s  "
              f"{bytes_per * n / med / 1e6:.0f} GB/s ({bytes_per} B/key)")

if "wide" in args.what.split(","):
    # glu_radix_sort_wide: 64-bit keys / wide payloads through the (key word, index) permutation
    sorter = glu.RadixSort()
    for name, kb, vb in [("u64 keys only", 8, 0), ("u64 keys + u32 values", 8, 4), ("u64 keys + u64 values", 8, 8),
                         ("u32 keys + 16-byte values", 4, 16)]:
        if kb == 8:
            k0 = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device=dev, generator=g) * 2
        else:
            k0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
        v0 = None if vb == 0 else torch.zeros((n, vb // 4), dtype=torch.int32, device=dev)
        k = k0.clone()
        v = None if v0 is None else v0.clone()

        def prep_w():
            k.copy_(k0)

        med, best = timeit(lambda: sorter.sort_wide(k, v, n, kb, vb), prep_w, reps=min(args.reps, 3))
        print(f"sort_wide {name:26s} n=2^{args.log2n}: median {med:.3f} ms  best {best:.3f} ms  {n / med / 1e6:.2f} Gkeys/s")
        del k0, v0, k, v

if "scan" in args.what:
    data0 = torch.randint(0, 100, (n,), dtype=torch.int32, device=dev, generator=g)
    data = data0.clone()
    scan = glu.BlellochScan(glu.DataType_Uint)
    med, best = timeit(lambda: scan(data, n), lambda: data.copy_(data0))
    print(f"scan   n=2^{args.log2n}: median {med:.3f} ms  best {best:.3f} ms  {8 * n / med / 1e6:.0f} GB/s")

if "reduce" in args.what:
    data0 = torch.randint(0, 100, (n,), dtype=torch.int32, device=dev, generator=g)
    data = data0.clone()
    red = glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)
    med, best = timeit(lambda: red(data, n), lambda: data.copy_(data0))
    print(f"reduce n=2^{args.log2n}: median {med:.3f} ms  best {best:.3f} ms  {4 * n / med / 1e6:.0f} GB/s")
    a = torch.empty(n, dtype=torch.int32, device=dev)
    med, best = timeit(lambda: a.copy_(data0))
    print(f"torch copy (read+write) n=2^{args.log2n}: median {med:.3f} ms  {8 * n / med / 1e6:.0f} GB/s")
    # what an IN-PLACE read-modify-write stream (the scan's access pattern: every line is read, then written) reaches
    med, best = timeit(lambda: a.add_(1))
    print(f"torch in-place add_ (read+write, same array) n=2^{args.log2n}: median {med:.3f} ms  {8 * n / med / 1e6:.0f} GB/s")

if "reducef" in args.what.split(","):
    # BASELINE.json configs[4]: Reduce(Float, Min/Max/Sum) over uniform floats in [-1, 1)
    f0 = (torch.rand(n, dtype=torch.float32, device=dev, generator=g) * 2 - 1)
    f = f0.clone()
    for name, op in [("Sum", glu.ReduceOperator_Sum), ("Min", glu.ReduceOperator_Min), ("Max", glu.ReduceOperator_Max)]:
        red = glu.Reduce(glu.DataType_Float, op)
        med, best = timeit(lambda: red(f, n), lambda: f.copy_(f0))
        want = {"Sum": f0.sum(dtype=torch.float64), "Min": f0.min(), "Max": f0.max()}[name].item()
        print(f"reduce Float {name} n=2^{args.log2n}: median {med:.3f} ms  best {best:.3f} ms  {4 * n / med / 1e6:.0f} GB/s  "
              f"result {f[0].item():.9g} (torch {want:.9g})")
