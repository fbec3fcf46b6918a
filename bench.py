#!/usr/bin/env python
"""bench.py — headline benchmark of the glu hot path on B200 (contract: see the task prompt / DESIGN.md §measurement).

Metric (BASELINE.json): Gpairs/s of RadixSort on 32-bit key + 32-bit value pairs.
  N = 1 : one step = one stable sort of 2^28 uniform-random uint32 (key, value) pairs (BASELINE configs[2]),
          inputs resident in HBM, K independent unsorted inputs (one per step; 2 GiB each, far larger than L2).
  N > 1 : weak scaling, 2^28 pairs per GPU per step, MSD split + NVLink all-to-all + local sort
          (gl-radix-sort_b200/distributed.py); value = pairs of all ranks / max-over-ranks device time.
          GLU_BENCH_MODE=pipeline (default: consecutive steps on two lanes, the exchange of step k+1 under the local
          sort of step k) or =serial (one DistributedRadixSort call after the other).  side_metrics carries the sharded
          Reduce / BlellochScan and BASELINE configs[3] (2^30 pairs per GPU).
Every line carries `verified`: the last timed step's output checked on the device (sorted, cross-rank order, multiset
checksum, stability).
Extra keys on the JSON line: roofline (dominant kernel = onesweep pass, 16 B/pair per launch, CUDA events on
the launch stream inside the timed region), cpu_baseline (std::stable_sort of the oracle on a bounded sample),
e2e (same sort through the host-buffer C-ABI entry point, pinned host memory, H2D + D2H inside), clocks,
gpu_launches, and the scan / reduce side metrics of BASELINE config 2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--log2-pairs 28]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SORT_BYTES_PER_PAIR = 68  # 4 B histogram read + 4 passes x (8 B read + 8 B write)   
PASS_BYTES_PER_PAIR = 16  # one onesweep launch: read key+value, write key+value


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--log2-pairs", type=int, default=28, help="pairs per GPU per step (log2)")
    p.add_argument("--cpu-sample-log2", type=int, default=28)
    p.add_argument("--no-side-metrics", action="store_true", help="skip scan/reduce/e2e/cpu legs (tuning runs)")
    p.add_argument("--config3-log2", type=int, default=30,
                   help="N > 1: pairs per GPU (log2) of the BASELINE configs[3] side measurement; <= --log2-pairs: off")
    p.add_argument("--config3-steps", type=int, default=5)
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    Sampled in-process through NVML (the library behind nvidia-smi; a query takes well under a millisecond, so even a
    ~50 ms timed region gets several samples — spawning `nvidia-smi -lms` per rank delivered its first line only after
    the run was over on an 8-GPU box).  Falls back to an nvidia-smi subprocess if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples = []   # (sm_mhz, max_mhz, set(reasons))
        self.proc = None
        self.thread = None
        self.running = False
        self.source = None

    # ---- NVML
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES may renumber the devices: go through the UUID
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)

    def _nvml_loop(self, nv, h):
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while self.running:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.samples.append((sm, mx, {name for bit, name in bits if mask & bit}))
                time.sleep(0.004)
        except Exception:
            pass

    # ---- nvidia-smi fallback
    def _smi_loop(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) < 9:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            self.samples.append((sm, mx, {name for name, v in zip(self.NAMES, parts[5:9])
                                          if v.lower().startswith("active")}))

    def start(self):
        try:
            nv, h = self._nvml_handle()
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.running = True
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.running = False
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """Number of samples seen so far (call at the start and at the end of the timed region)."""
        return len(self.samples)

    def stop(self, first=0, last=None):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        if self.source == "nvidia-smi" and not self.samples:
            for _ in range(40):  # its first line can take seconds on a multi-GPU box
                if self.samples:
                    break
                time.sleep(0.1)
        time.sleep(0.02)
        self.running = False
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        # samples taken inside the timed region (one before and one after included)
        window = self.samples[max(0, first - 1):(last + 1 if last is not None else None)] or self.samples[-3:]
        sm = sorted(x[0] for x in window)
        reasons = set()
        for x in window:
            reasons |= x[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": window[-1][1] if window else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------------ shared pieces

def host_threads() -> int:
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm sizes its
    thread team from the affinity mask instead (the oracle sets the team size explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(args) -> dict:
    """`config` of the JSON line — it names the WORKLOAD and is identical on both arms (how each arm runs it is in the
    line's `implementation` object)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"workload": f"RadixSort of 2^{args.log2_pairs} uniform-random uint32 key/value pairs per GPU "
                        f"(BASELINE.json configs[2]), values = input index, one fresh unsorted input per step",
            "pairs_per_gpu": 1 << args.log2_pairs,
            "parallelism": f"{world} shard(s) of 2^{args.log2_pairs} pairs, one per GPU (weak scaling); the result is "
                           f"the stable sort of the concatenated shards, rank r's slice before rank r+1's",
            "cache": f"every step sorts a fresh unsorted input of {8 << args.log2_pairs >> 20} MiB per shard - larger "
                     f"than the 126 MB L2 (and any host cache): no flush needed"}


def gl_probe() -> str:
    """north_star's second CPU comparator — the reference's GLSL shaders on Mesa llvmpipe — needs an OpenGL 4.6
    context on the host.  Probe for the libraries instead of assuming."""
    import ctypes.util

    found = [n for n in ("GL", "EGL", "OSMesa") if ctypes.util.find_library(n)]
    if not found:
        return "unavailable (no GL 4.6 on host: libGL / libEGL / libOSMesa not found)"
    return f"unavailable (found {'/'.join('lib' + n for n in found)} but the reference's GLFW + glad build is not part of this repo)"


# ------------------------------------------------------------------------------------------------ reference arm

def run_reference(args):
    """The reference's own CPU path for this metric: the std::stable_sort of (key,value) pairs its test-suite
    oracle prescribes (test/radix_sort_tests.cpp:20-51 strengthened per north_star; oracle/glu_oracle.cpp),
    on all host threads (__gnu_parallel::stable_sort).  The reference's GPU path is GLSL on an OpenGL 4.6
    context and cannot run on this box (no GL/X11), so oracle/_ref does not exist — kind = "port".
    One step = one sort of a bounded 2^cpu_sample_log2-pair sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    import oracle

    oracle.build()
    n = 1 << args.cpu_sample_log2
    threads = host_threads()
    keys = oracle.mt19937_u32(1, n)
    vals = np.arange(n, dtype=np.uint32)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget_s = 240.0  # wall-clock bound for the whole arm (warm-up included)
    t_begin = time.time()
    done_warmup = 0
    for _ in range(warmup):
        oracle.time_stable_sort_pairs(keys, vals, threads)
        done_warmup += 1
        if time.time() - t_begin > budget_s / 4:
            break
    times = []
    for _ in range(steps):
        times.append(oracle.time_stable_sort_pairs(keys, vals, threads))
        if time.time() - t_begin > budget_s:
            break
    total = sum(times)
    value = n * len(times) / total / 1e9
    cfg = workload_config(args)
    implementation = {
        "parallelism": f"host CPU, {threads} threads (__gnu_parallel::stable_sort)",
        "cache": "each step sorts a fresh copy of the sample (2 GiB, larger than any host cache)",
        "reference_arm": (f"std::stable_sort of the (key, value) pairs on the host (the reference test-suite's oracle; "
                          f"its GLSL path needs OpenGL 4.6), each step a 2^{args.cpu_sample_log2}-pair uniform-random "
                          f"(mt19937) sample of that workload")}
    line = {
        "impl": "reference", "metric": "radix_sort_u32_key_value_throughput", "value": value, "unit": "Gpairs/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": done_warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": cfg, "implementation": implementation,
        "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                         "sample": f"2^{args.cpu_sample_log2} pairs per step, __gnu_parallel::stable_sort, "
                                   f"{threads} threads"},
        "glsl_llvmpipe": gl_probe(),
        "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if len(times) < steps:
        line["note"] = f"stopped after {len(times)} of {steps} steps: {budget_s:.0f} s wall-clock budget"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm

def cpu_baseline(args):
    import numpy as np

    import oracle

    oracle.build()
    n = 1 << args.cpu_sample_log2
    threads = host_threads()
    keys = oracle.mt19937_u32(1, n)
    vals = np.arange(n, dtype=np.uint32)
    t = oracle.time_stable_sort_pairs(keys, vals, threads)
    return {"value": n / t / 1e9, "unit": "Gpairs/s", "cores": threads, "kind": "port",
            "sample": f"one __gnu_parallel::stable_sort of the first 2^{args.cpu_sample_log2} pairs "
                      f"(mt19937 keys, index values), {threads} threads, {t:.2f} s",
            "glsl_llvmpipe": gl_probe()}


def pinned_u32(glu, n):
    """A pinned host uint32 array of n elements (cudaMallocHost through the C ABI)."""
    import numpy as np

    ptr = ctypes.c_void_p()
    glu.check(glu.lib.glu_malloc_host(ctypes.byref(ptr), 4 * n), "glu_malloc_host")
    buf = (ctypes.c_uint32 * n).from_address(ptr.value)
    return np.frombuffer(buf, dtype=np.uint32), ptr


def side_metrics(glu, torch, dev, n, peak):
    """BASELINE config 2: Reduce(Uint, Sum) and BlellochScan(Uint) over 2^28 uint32, GB/s vs HBM."""
    out = {}
    g = torch.Generator(device=dev).manual_seed(7)
    data0 = torch.randint(0, 100, (n,), dtype=torch.int32, device=dev, generator=g)
    data = data0.clone()
    for name, op, kid, bytes_per_elem in (("scan", glu.BlellochScan(glu.DataType_Uint), glu.KERNEL_SCAN, 8),
                                          ("reduce", glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum),
                                           glu.KERNEL_REDUCE, 4)):
        for _ in range(3):
            data.copy_(data0)
            op(data, n)
        torch.cuda.synchronize()
        glu.profile_collect(kid)
        reps = 10
        for _ in range(reps):
            data.copy_(data0)  # also evicts the previous result from L2 (2 GiB of traffic > 126 MB L2)
            op(data, n)
        torch.cuda.synchronize()
        ms, launches = glu.profile_collect(kid)
        gbs = bytes_per_elem * n * launches / ms / 1e6
        out[name] = {"n": n, "ms": ms / launches, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak,
                     "bytes_per_elem": bytes_per_elem}
    return out


def distributed_side_metrics(glu, torch, dist, dev, world, rank, n, peak):
    """Multi-GPU Reduce(Uint, Sum) and BlellochScan(Uint), 2^28 elements per GPU, sharded by contiguous ranges
    (glu/Reduce.hpp:111-135 and glu/BlellochScan.hpp:130-139 semantics over the concatenation of the shards).
    Device time of `reps` back-to-back calls between barriers, max over ranks; every call restores its input first
    (a 1 GiB device copy, excluded by timing the copies alone and subtracting them)."""
    out = {}
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    data0 = torch.randint(0, 100, (n,), dtype=torch.int32, device=dev, generator=g)
    data = data0.clone()
    expect_total = torch.tensor([int(data0.sum(dtype=torch.int64).item())], dtype=torch.int64, device=dev)
    totals = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(totals, expect_total)
    totals = [int(t.item()) for t in totals]
    reps = 10

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    copy_ms = timed(lambda: data.copy_(data0))
    red = glu.DistributedReduce(glu.DataType_Uint, glu.ReduceOperator_Sum)
    scan = glu.DistributedBlellochScan(glu.DataType_Uint)

    def do_reduce():
        data.copy_(data0)
        red(data, n)

    def do_scan():
        data.copy_(data0)
        scan(data, n)

    for name, fn, bytes_per_elem, note in (
            ("reduce", do_reduce, 4, "local reduce kernel + all-gather of one element per rank + P-element combine"),
            ("scan", do_scan, 12, "local reduce for the rank total (4 B) + all-gather + seeded single-pass scan (8 B); "
                                  "a variant keeping each tile's inclusive total from the scan itself would move 8 B")):
        ms = max(1e-6, timed(fn) - copy_ms)
        gbs = bytes_per_elem * n * world / ms / 1e6
        out[name] = {"n_per_gpu": n, "ms": ms, "GB/s_total": gbs, "frac_of_N_x_hbm_peak": gbs / (peak * world),
                     "bytes_per_elem": bytes_per_elem, "how": note}
        if name == "scan":  # algorithmic 8 B/elem view of the same time
            out[name]["GB/s_total_at_8B"] = 8 * n * world / ms / 1e6
            out[name]["frac_at_8B"] = 8 * n * world / ms / 1e6 / (peak * world)
    # parity of the last calls (cheap, on device): reduce result = sum of all shards mod 2^32; scan's first element on
    # rank r = sum of the shards before it, its last element = that + its own total - its last input
    do_reduce()
    got = int(data[0].item()) & 0xFFFFFFFF
    assert got == sum(totals) & 0xFFFFFFFF, ("DistributedReduce mismatch", got, sum(totals) & 0xFFFFFFFF)
    do_scan()
    first, last = int(data[0].item()) & 0xFFFFFFFF, int(data[n - 1].item()) & 0xFFFFFFFF
    base = sum(totals[:rank]) & 0xFFFFFFFF
    assert first == base, ("DistributedBlellochScan base mismatch", first, base)
    assert last == (base + totals[rank] - int(data0[n - 1].item())) & 0xFFFFFFFF, "DistributedBlellochScan last element"
    out["verified"] = True
    return out


# ---- on-device verification of a (possibly sharded) sort result: SURVEY.md §7 last bullet

_MIX1 = 0x9E3779B97F4A7C15 - (1 << 64)
_MIX2 = 0xBF58476D1CE4E5B9 - (1 << 64)
_CHUNK = 1 << 25


def _ordered(x):
    """int32 tensor holding uint32 bit patterns -> int32 tensor whose SIGNED order is the unsigned order."""
    return x ^ (-2147483648)


def multiset_checksum(torch, keys, vals, count):
    """Order-independent 64-bit checksum of the (key, value) multiset: sum over pairs of mix64(key << 32 | value)."""
    acc = torch.zeros((), dtype=torch.int64, device=keys.device)
    for s in range(0, count, _CHUNK):
        e = min(count, s + _CHUNK)
        x = (keys[s:e].to(torch.int64) << 32) | (vals[s:e].to(torch.int64) & 0xFFFFFFFF)
        x = x * _MIX1
        x = x ^ (x >> 29)
        x = x * _MIX2
        x = x ^ (x >> 32)
        acc = acc + x.sum()
    return acc


def verify_sorted(torch, dist, world, rank, dev, out_keys, out_vals, m, in_checksum, total_pairs):
    """Checks on the device, with a handful of scalars exchanged between ranks:
    sorted (every rank's slice non-decreasing), rank_boundaries (last pair of rank r <= first pair of rank r + 1),
    multiset_checksum (the output holds exactly the input pairs), stable (values non-decreasing inside every run of
    equal keys — the input values increase with the global input position), pairs (nothing lost or duplicated)."""
    sorted_ok = stable_ok = True
    for s in range(0, max(m - 1, 0), _CHUNK):
        e = min(m - 1, s + _CHUNK)
        k = _ordered(out_keys[s:e + 1])
        v = _ordered(out_vals[s:e + 1])
        sorted_ok = sorted_ok and bool((k[1:] >= k[:-1]).all())
        stable_ok = stable_ok and not bool(((k[1:] == k[:-1]) & (v[1:] < v[:-1])).any())
    out_checksum = multiset_checksum(torch, out_keys, out_vals, m)
    edge = torch.zeros(5, dtype=torch.int64, device=dev)
    if m > 0:
        edge[0] = out_keys[0].to(torch.int64) & 0xFFFFFFFF
        edge[1] = out_vals[0].to(torch.int64) & 0xFFFFFFFF
        edge[2] = out_keys[m - 1].to(torch.int64) & 0xFFFFFFFF
        edge[3] = out_vals[m - 1].to(torch.int64) & 0xFFFFFFFF
    edge[4] = m
    sums = torch.stack([in_checksum, out_checksum, torch.tensor(int(sorted_ok and stable_ok), dtype=torch.int64, device=dev),
                        torch.tensor(int(sorted_ok), dtype=torch.int64, device=dev)])
    if world > 1:
        edges = [torch.zeros(5, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(edges, edge)
        gathered = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, sums)
    else:
        edges, gathered = [edge], [sums]
    edges = [[int(x) for x in e.tolist()] for e in edges]
    gathered = [[int(x) for x in g.tolist()] for g in gathered]
    mask = (1 << 64) - 1
    in_sum = sum(g[0] for g in gathered) & mask
    out_sum = sum(g[1] for g in gathered) & mask
    boundaries_ok = True
    prev = None
    for e in edges:
        if e[4] == 0:
            continue
        if prev is not None and not (prev[2] < e[0] or (prev[2] == e[0] and prev[3] <= e[1])):
            boundaries_ok = False
        prev = e
    pairs = sum(e[4] for e in edges)
    result = {"sorted": all(g[3] == 1 for g in gathered), "rank_boundaries": boundaries_ok,
              "multiset_checksum": in_sum == out_sum, "stable": all(g[2] == 1 for g in gathered) and boundaries_ok,
              "pairs": pairs, "pairs_expected": total_pairs,
              "how": "on device, last timed step: every rank's full output non-decreasing; last pair of rank r <= first "
                     "pair of rank r+1; sum of mix64(key,value) over inputs == over outputs (all-gathered); values "
                     "non-decreasing inside equal-key runs (input values increase with global input position)"}
    result["ok"] = bool(result["sorted"] and result["rank_boundaries"] and result["multiset_checksum"]
                        and result["stable"] and pairs == total_pairs)
    return result


class InputRing:
    """`slots` independent unsorted (keys, values) inputs resident in HBM.  Keys: uniform random 32-bit patterns
    (torch's Philox generator, seeded per rank); values: the pair's global input position (>> vshift when the job holds
    more than 2^32 pairs), so a stable sort leaves the values of equal keys in increasing order."""

    def __init__(self, torch, dev, n, slots, rank, world, seed=1):
        self.torch, self.dev, self.n, self.rank = torch, dev, n, rank
        self.gen = torch.Generator(device=dev).manual_seed(seed + rank)
        total = world * n
        self.vshift = max(0, (total - 1).bit_length() - 32)
        self.slots = [(torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev))
                      for _ in range(slots)]
        for i in range(slots):
            self.refill(i)

    def refill(self, i):
        torch, n = self.torch, self.n
        k, v = self.slots[i]
        step = 1 << 26
        for s in range(0, n, step):
            e = min(n, s + step)
            k[s:e] = torch.randint(-(1 << 31), (1 << 31) - 1, (e - s,), dtype=torch.int32, device=self.dev,
                                   generator=self.gen)
            g = (torch.arange(s, e, dtype=torch.int64, device=self.dev) + self.rank * n) >> self.vshift
            v[s:e] = (((g + (1 << 31)) & 0xFFFFFFFF) - (1 << 31)).to(torch.int32)  # uint32 bit pattern
        return k, v


def describe_exchange(sorter) -> str:
    """What a DistributedRadixSort does between the plan and the result, in words (for the line's `config`)."""
    if sorter.exchange != "p2p":
        return "partition + NCCL all_to_all, local onesweep sort on all 32 bits"
    if sorter.local != "segmented":
        return "fused partition + NVLink peer-store all-to-all, local onesweep sort on all 32 bits"
    how = {"dma": "local MSD pass into tile-aligned staging + one copy-engine transfer per peer over NVLink (host plan)",
           "staged": "local MSD pass into staging + NVLink peer-store copy kernel",
           "direct": "MSD pass storing straight into the peers' memory"}[sorter.exchange_style]
    return f"{how}, segmented local onesweep sort on the 24 key bits below the split digit"


def measure_sort(args, glu, torch, dist, dev, world, rank, n, steps, warmup, mode, peak, input_budget_bytes):
    """K timed steps of the sort at `n` pairs per GPU.  Returns (line fields, objects to keep alive / close)."""
    in_place = world == 1  # glu::RadixSort sorts the caller's arrays; the multi-GPU sort leaves its input alone
    if in_place:
        slots = max(2, min(steps + warmup, input_budget_bytes // (8 * n)))
    else:
        slots = 2
    ring = InputRing(torch, dev, n, slots, rank, world)
    last = {}
    closers = []
    if world > 1:
        if mode == "pipeline":
            pipe = glu.DistributedSortPipeline(n)
            closers.append(pipe.close)
            exchange = "p2p"
            last["local"] = pipe.lanes[0].local

            def submit(k, v):
                last["ticket"] = pipe.submit(k, v, n)

            def finish():
                pipe.flush()

            def result():
                return pipe.result(last["ticket"])

            parallelism = (f"msd-split x{world}: top-8-bit histogram all-gather, balanced bucket->GPU prefix, "
                           f"{describe_exchange(pipe.lanes[0])}; consecutive steps software-pipelined on "
                           f"{pipe.num_lanes} lanes (the NVLink-bound exchange of step k+1 runs under the local sort of step k)")
        else:
            dsort = glu.DistributedRadixSort(n, exchange=os.environ.get("GLU_BENCH_EXCHANGE", "auto"))
            closers.append(dsort.close)
            exchange = dsort.exchange
            last["local"] = dsort.local

            def submit(k, v):
                last["out"] = dsort(k, v, n)

            def finish():
                pass

            def result():
                return last["out"]

            parallelism = (f"msd-split x{world}: top-8-bit histogram all-gather, balanced bucket->GPU prefix, "
                           f"{describe_exchange(dsort)}; steps back to back, not overlapped")
    else:
        sorter = glu.RadixSort()
        sorter.prepare_internal_buffers(n)  # as the reference's benchmark does (test/radix_sort_tests.cpp:187)

        def submit(k, v):
            sorter(k, v, n)
            last["out"] = (k, v, n)

        def finish():
            pass

        def result():
            return last["out"]

        parallelism = "single GPU"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed).  In-place sorting consumes inputs: refill what the warm-up used.
    used = 0
    for i in range(warmup):
        if in_place and used == slots:
            barrier()
            for j in range(slots):
                ring.refill(j)
            used = 0
        submit(*ring.slots[used % slots])
        used += 1
    finish()
    barrier()
    if in_place and used + min(steps, slots) > slots:
        for j in range(used):
            ring.refill(j)
        used = 0
    glu.profile_enable(True)
    for kid in (glu.KERNEL_SORT_ONESWEEP, glu.KERNEL_SORT_HISTOGRAM, glu.KERNEL_SORT_PARTITION):
        glu.profile_collect(kid)
    launches0 = glu.kernel_launch_count()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()

    # ---- timed: chunks of back-to-back steps, each chunk bracketed by barrier + synchronize and a CUDA event pair.
    # One chunk holds all K steps whenever K fresh inputs fit in HBM (the default); otherwise inputs are regenerated
    # on the device between chunks, outside the events.
    ms_total, done, chunks = 0.0, 0, 0
    mark0 = None
    in_checksum = None
    while done < steps:
        c = steps - done if not in_place else min(steps - done, slots - used)
        if c == 0:
            for j in range(slots):
                ring.refill(j)
            used = 0
            continue
        if done + c == steps:  # the chunk holding the last step: checksum of that step's input, before it is sorted
            k_last, v_last = ring.slots[(used + c - 1) % slots]
            in_checksum = multiset_checksum(torch, k_last, v_last, n)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if mark0 is None:
            mark0 = sampler.mark()
        ev0.record()
        for i in range(c):
            submit(*ring.slots[(used + i) % slots])
        finish()
        ev1.record()
        barrier()
        ms_total += ev0.elapsed_time(ev1)
        used += c if in_place else 0
        done += c
        chunks += 1
    clocks = sampler.stop(mark0, sampler.mark())
    gpu_launches = glu.kernel_launch_count() - launches0
    sweep_ms, sweep_launches = glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
    hist_ms, hist_launches = glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
    part_ms, part_launches = glu.profile_collect(glu.KERNEL_SORT_PARTITION)
    glu.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

    # ---- the last timed step really sorted its input (all ranks take part)
    ok_, ov_, m = result()
    verified = verify_sorted(torch, dist, world, rank, dev, ok_, ov_, int(m), in_checksum, world * n)
    assert verified["ok"], f"bench output failed verification: {verified}"

    value = world * n * steps / (ms_total * 1e-3) / 1e9
    # HBM traffic of the SM kernels per pair and step.  One GPU: 68 B.  N > 1, segmented local sort: split-digit histogram
    # 4 + MSD pass 16 + segment histograms 4 + 3 local passes 48 = 72 B (north_star's figure; the copy engines move
    # another ~14 B per pair between staging and the peers' receive arrays).  N > 1, full local sort: 4 + 16 + 68 = 88 B.
    segmented = world > 1 and last.get("local") == "segmented"
    step_bytes = SORT_BYTES_PER_PAIR if world == 1 else (72 if segmented else SORT_BYTES_PER_PAIR + 4 + PASS_BYTES_PER_PAIR)
    per_launch_ms = sweep_ms / max(1, sweep_launches)
    # single GPU: every launch sweeps the whole array; N > 1: the local sort sweeps what the rank received (~n)
    achieved = PASS_BYTES_PER_PAIR * n / (per_launch_ms * 1e-3) / 1e9 if sweep_launches else None
    roofline = {"bound": "hbm", "kernel": "onesweep pass (one 8-bit digit; onesweep_kernel<320,24,3,Ballot>, the SEG flavour in the multi-GPU local sort)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": None, "launches": sweep_launches, "ms_per_launch": per_launch_ms,
                "algorithmic_bytes_per_launch": PASS_BYTES_PER_PAIR * n,
                "kernel_share_of_step": sweep_ms / ms_total,
                "histogram_ms_per_launch": hist_ms / max(1, hist_launches),
                "partition_exchange_ms_per_launch": (part_ms / part_launches) if part_launches else None,
                # per GPU and pair: the local sort's 68 B, plus at N > 1 the split-digit histogram (4 B) and the
                # partition pass (8 B read, 8 B written to local or peer memory)
                "whole_sort": {"bytes_per_pair": step_bytes,
                               "achieved_GB/s": step_bytes * n * steps / (ms_total * 1e-3) / 1e9,
                               "frac": step_bytes * n * steps / (ms_total * 1e-3) / 1e9 / peak}}
    if world > 1 and mode == "pipeline":
        roofline["note"] = ("the lanes overlap: per-launch times are those of kernels sharing the GPU with the other "
                            "lane's kernels, so achieved is a lower bound of the kernel alone")
    fields = {"value": value, "ms_per_step": ms_total / steps, "roofline": roofline, "clocks": clocks,
              "gpu_launches": int(gpu_launches), "verified": verified, "parallelism": parallelism,
              "timed_chunks": chunks, "input_slots": slots}
    keep = {"ring": ring, "closers": closers}
    return fields, keep


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    glu = entry.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << args.log2_pairs
    steps, warmup = args.steps, max(3, args.warmup)
    peak, peak_src = measured_peaks()
    mode = os.environ.get("GLU_BENCH_MODE", "pipeline" if world > 1 else "single")
    total_mem = torch.cuda.get_device_properties(dev).total_memory
    input_budget = int(total_mem * 0.40)  # inputs may take 40 % of HBM; the rest is scratch, receive lanes, verification

    fields, keep = measure_sort(args, glu, torch, dist, dev, world, rank, n, steps, warmup, mode, peak, input_budget)
    roofline = fields.pop("roofline")
    roofline["peak_source"] = peak_src
    traffic_file = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass
    parallelism = fields.pop("parallelism")
    cache = (f"inputs ({8 * n >> 20} MiB per step) are larger than the 126 MB L2; no flush needed; "
             f"{fields.pop('input_slots')} inputs resident, {fields.pop('timed_chunks')} timed chunk(s)")
    line = {
        "metric": "radix_sort_u32_key_value_throughput", "value": fields["value"], "unit": "Gpairs/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": fields["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args), "implementation": {"parallelism": parallelism, "cache": cache},
        "roofline": roofline, "clocks": fields["clocks"], "gpu_launches": fields["gpu_launches"],
        "verified": fields["verified"],
        "published_reference": {"value": 0.05345, "unit": "Gpairs/s", "hardware": "RTX 2060 SUPER (README.md:133)",
                                "note": "different hardware; not used for vs_baseline"},
    }
    ring = keep["ring"]

    if rank == 0 and world == 1 and not args.no_side_metrics:
        del ring.slots[1:]
        torch.cuda.empty_cache()
        line["e2e"] = e2e_single(args, glu, torch, np, n, steps)
        glu.profile_enable(True)
        line["side_metrics"] = side_metrics(glu, torch, dev, 1 << 28, peak)
        glu.profile_enable(False)
        line["cpu_baseline"] = cpu_baseline(args)
    elif world > 1 and not args.no_side_metrics:
        for close in keep["closers"]:
            close()
        keep["closers"] = []
        del ring.slots[1:]
        torch.cuda.empty_cache()
        line["e2e"] = e2e_distributed(args, glu, torch, dist, dev, world, rank, n, steps, ring.slots[0])
        del ring, keep
        torch.cuda.empty_cache()
        line["side_metrics"] = distributed_side_metrics(glu, torch, dist, dev, world, rank, 1 << 28, peak)
        # ---- BASELINE.json configs[3]: 2^30 pairs per GPU (2^31 / 2^32 / 2^33 pairs in the job)
        if args.config3_log2 > args.log2_pairs:
            torch.cuda.empty_cache()
            c3_steps = max(2, min(steps, args.config3_steps))
            f3, k3 = measure_sort(args, glu, torch, dist, dev, world, rank, 1 << args.config3_log2, c3_steps, 3, mode, peak,
                                  input_budget)
            for close in k3["closers"]:
                close()
            line["side_metrics"]["config3"] = {
                "workload": f"RadixSort of 2^{args.config3_log2} pairs per GPU, {world} GPUs: "
                            f"{world << args.config3_log2} pairs per step (BASELINE.json configs[3])",
                "value": f3["value"], "unit": "Gpairs/s", "steps": c3_steps, "warmup": 3, "ms_per_step": f3["ms_per_step"],
                "verified": f3["verified"], "clocks": f3["clocks"],
                "onesweep_ms_per_launch": f3["roofline"]["ms_per_launch"],
                "partition_exchange_ms_per_launch": f3["roofline"]["partition_exchange_ms_per_launch"]}
            del k3
    elif rank == 0:
        line["e2e"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_single(args, glu, torch, np, n, steps):
    """e2e at one GPU: the same sort through the host-buffer C-ABI, pinned host memory, copies inside the timed region.
    The steps go through glu_host_sort_queue (depth 3: with two device slots the upload of step k+2 would wait for the
    download of step k on the same slot, i.e. the upload engine would idle for one sort per step): the upload of step
    k+1 overlaps the sort and the download of step k, every step still pays its own H2D + sort + D2H.  Results land in
    place, so every step needs its own unsorted pinned input: all K inputs are pinned at once when the host has the
    memory for it (a quarter of MemAvailable), otherwise the steps are timed in chunks of 8 inputs refilled between
    chunks outside the clock (every chunk then pays the queue's fill and drain).  The synchronous single call
    (glu_radix_sort_u32kv_host) is timed beside it."""
    slots = min(steps, 8)
    try:
        avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        if 8 * n * min(steps, 24) <= avail // 4:
            slots = min(steps, 24)
    except Exception:
        pass
    host = []
    for i in range(slots):
        hk, hk_ptr = pinned_u32(glu, n)
        hv, hv_ptr = pinned_u32(glu, n)
        host.append((hk, hv, hk_ptr, hv_ptr))
    seed = [100]

    def refill(count):
        for hk, hv, _, _ in host[:count]:
            hk[:] = np.random.default_rng(seed[0]).integers(0, 1 << 32, size=n, dtype=np.uint32)
            hv[:] = np.arange(n, dtype=np.uint32)
            seed[0] += 1

    queue = glu.HostSortQueue(n, depth=3)
    wk, wk_ptr = pinned_u32(glu, 1 << 20)
    wv, wv_ptr = pinned_u32(glu, 1 << 20)
    wk[:] = np.arange(1 << 20, dtype=np.uint32)[::-1]
    wv[:] = 0
    for _ in range(3):
        queue.submit(wk, wv)  # warms the queue's streams
    queue.wait()
    e2e_t, done = 0.0, 0
    while done < steps:
        c = min(slots, steps - done)
        refill(c)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for hk, hv, _, _ in host[:c]:
            queue.submit(hk, hv, n)
        queue.wait()
        e2e_t += time.perf_counter() - t0
        done += c
        for hk, hv, _, _ in host[:c]:
            assert bool(np.all(hk[:-1][: 1 << 22] <= hk[1:][: 1 << 22])), "e2e output is not sorted"
    queue.close()
    # synchronous call on a fresh input (first call warms its allocation path)
    glu.radix_sort_u32kv_host(wk, wv)
    hk, hv = host[0][0], host[0][1]
    hk[:] = np.random.default_rng(99).integers(0, 1 << 32, size=n, dtype=np.uint32)
    hv[:] = np.arange(n, dtype=np.uint32)
    glu.radix_sort_u32kv_host(hk, hv, n)
    hk[:] = np.random.default_rng(98).integers(0, 1 << 32, size=n, dtype=np.uint32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    glu.radix_sort_u32kv_host(hk, hv, n)
    sync_t = time.perf_counter() - t0
    out = {"value": n * steps / e2e_t / 1e9, "unit": "Gpairs/s", "h2d_bytes_per_step": 8 * n,
           "d2h_bytes_per_step": 8 * n, "steps": steps, "ms_per_step": 1e3 * e2e_t / steps,
           "api": "glu_host_sort_queue (depth 3; per step: pinned host arrays -> H2D -> sort -> D2H, "
                  f"consecutive steps overlapped; {slots} pinned inputs per timed chunk)",
           "synchronous_call": {"value": n / sync_t / 1e9, "ms": 1e3 * sync_t, "api": "glu_radix_sort_u32kv_host"}}
    for _, _, p0, p1 in host:
        glu.lib.glu_free_host(p0)
        glu.lib.glu_free_host(p1)
    glu.lib.glu_free_host(wk_ptr)
    glu.lib.glu_free_host(wv_ptr)
    return out


def e2e_distributed(args, glu, torch, dist, dev, world, rank, n, steps, base):
    """e2e at N GPUs: every rank uploads its shard from pinned host memory, the job sorts, every rank downloads its slice
    of the result.  The sorts go through the two-lane pipeline (DistributedSortPipeline): the result of step k stays in
    its lane until step k+2 is submitted, so its download runs on a copy stream of its own while step k+1 is exchanged
    and sorted, and the upload of step k+1 (second copy stream, second device input) runs under both — PCIe is full
    duplex.  Every step still pays its own H2D + sort + D2H inside the timed region (wall clock between barriers, max
    over ranks)."""
    e2e_steps = max(2, min(steps, 8))
    base_k, base_v = base[0].cpu(), base[1].cpu()
    pipe = None
    try:  # per rank at 2^28 pairs: 3 + 2 x 1.25 GiB x 2 of pinned host memory, 4 GiB of HBM for the inputs
        pipe = glu.DistributedSortPipeline(n)
        cap = pipe.lanes[0].capacity
        hks = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]  # two unsorted shards, alternating
        hv = torch.empty(n, dtype=torch.int32).pin_memory()
        hv.copy_(base_v)
        for i, hk in enumerate(hks):
            hk.copy_(base_k ^ (0x9E3779B9 * (i + 1) & 0x7FFFFFFF))
        oks = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        ovs = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
        dks = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
        dvs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
        allocated = 1
    except RuntimeError as e:  # e.g. the host cannot pin that much for every rank
        allocated = 0
        sys.stderr.write(f"[bench rank {rank}] e2e buffers: {e}\n")
    flag = torch.tensor([allocated], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks run the e2e leg or none does (it contains collectives)
    if int(flag.item()) == 0:
        if pipe is not None:
            pipe.close()
        return {"value": None, "skipped": "a rank could not allocate the pinned host / device buffers of the e2e leg"}
    up_stream = torch.cuda.Stream(device=dev)
    down_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]    # the exchange that read device input i % 2 has been enqueued
    downloaded = [torch.cuda.Event(), torch.cuda.Event()]  # lane i % 2's result has left the device

    def upload(i):
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(consumed[i % 2])
            dks[i % 2].copy_(hks[i % 2], non_blocking=True)
            dvs[i % 2].copy_(hv, non_blocking=True)
            uploaded[i % 2].record(up_stream)

    last_m = [0]

    def run(first, count):
        upload(first)
        tickets = []
        for i in range(first, first + count):
            main_stream.wait_event(uploaded[i % 2])
            main_stream.wait_event(downloaded[i % 2])  # lane i % 2 is about to be overwritten by this step's exchange
            if i + 1 < first + count:
                upload(i + 1)
            t = pipe.submit(dks[i % 2], dvs[i % 2], n)
            # the partition pass of this submit is the last reader of the device input
            consumed[i % 2].record(pipe.stream_x)
            tickets.append((i, t))
            if len(tickets) == 2:
                download(*tickets.pop(0))
        while tickets:
            download(*tickets.pop(0))

    def download(i, t):
        with torch.cuda.stream(down_stream):
            sk, sv, m = pipe.result(t)  # makes down_stream wait for the job's local sort
            oks[i % 2][:m].copy_(sk, non_blocking=True)
            ovs[i % 2][:m].copy_(sv, non_blocking=True)
            downloaded[i % 2].record(down_stream)
            last_m[0] = m

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for ev in consumed + downloaded:
        ev.record(main_stream)
    run(0, 2)  # warm-up steps (pinned-copy paths, the copy streams, both lanes)
    barrier()
    t0 = time.perf_counter()
    run(2, e2e_steps)
    barrier()
    e2e_t = time.perf_counter() - t0
    i_last = (2 + e2e_steps - 1) % 2
    m = last_m[0]
    out64 = _ordered(oks[i_last][: min(m, 1 << 22)])
    assert bool((out64[1:] >= out64[:-1]).all()), "e2e output is not sorted"
    t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_t = float(t.item())
    pipe.close()
    return {"value": world * n * e2e_steps / e2e_t / 1e9, "unit": "Gpairs/s",
            "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": 8 * n * world, "steps": e2e_steps,
            "ms_per_step": 1e3 * e2e_t / e2e_steps,
            "api": "DistributedSortPipeline (per rank and step: pinned host shard -> H2D -> sort -> D2H of its slice; "
                   "uploads, sorts and downloads of consecutive steps overlap on three streams and two receive lanes)"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
