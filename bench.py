#!/usr/bin/env python
"""bench.py — headline benchmark of the glu hot path on B200 (contract: see the task prompt / DESIGN.md §measurement).

Metric (BASELINE.json): Gpairs/s of RadixSort on 32-bit key + 32-bit value pairs.
  N = 1 : one step = one stable sort of 2^28 uniform-random uint32 (key, value) pairs (BASELINE configs[2]),
          inputs resident in HBM, K independent unsorted inputs (one per step; 2 GiB each, far larger than L2).
  N > 1 : weak scaling, 2^28 pairs per GPU per step, MSD split + NVLink all-to-all + local sort
          (gl-radix-sort_b200/distributed.py); value = pairs of all ranks / max-over-ranks device time.
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
    threads = oracle.max_threads()
    keys = oracle.mt19937_u32(1, n)
    vals = np.arange(n, dtype=np.uint32)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget_s = 150.0
    t_begin = time.time()
    for _ in range(min(warmup, 1)):
        oracle.time_stable_sort_pairs(keys, vals, threads)
    times = []
    for _ in range(steps):
        times.append(oracle.time_stable_sort_pairs(keys, vals, threads))
        if time.time() - t_begin > budget_s:
            break
    total = sum(times)
    value = n * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "radix_sort_u32_key_value_throughput", "value": value, "unit": "Gpairs/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(warmup, 1), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"RadixSort of 2^{args.log2_pairs} uniform-random uint32 key/value pairs per GPU "
                               f"(BASELINE.json configs[2]), values = input index, one fresh unsorted input per step",
                   "pairs_per_gpu": 1 << args.log2_pairs,
                   "reference_arm": f"std::stable_sort of the (key, value) pairs on the host (the reference test-suite's "
                                    f"oracle; its GLSL path needs OpenGL 4.6), each step a 2^{args.cpu_sample_log2}-pair "
                                    f"uniform-random (mt19937) sample of that workload"},
        "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                         "sample": f"2^{args.cpu_sample_log2} pairs per step, __gnu_parallel::stable_sort, "
                                   f"{threads} threads"},
        "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm

def cpu_baseline(args):
    import numpy as np

    import oracle

    oracle.build()
    n = 1 << args.cpu_sample_log2
    threads = oracle.max_threads()
    keys = oracle.mt19937_u32(1, n)
    vals = np.arange(n, dtype=np.uint32)
    t = oracle.time_stable_sort_pairs(keys, vals, threads)
    return {"value": n / t / 1e9, "unit": "Gpairs/s", "cores": threads, "kind": "port",
            "sample": f"one __gnu_parallel::stable_sort of the first 2^{args.cpu_sample_log2} pairs "
                      f"(mt19937 keys, index values), {threads} threads, {t:.2f} s"}


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

    # ---- inputs: one independent unsorted (keys, vals) pair per step, resident in HBM before timing starts
    total_inputs = steps + warmup
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    inputs = []
    for _ in range(total_inputs):
        k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=gen)
        v = torch.arange(n, dtype=torch.int32, device=dev)
        inputs.append((k, v))

    last_out = [None]
    if world > 1:
        dsort = glu.DistributedRadixSort(n, exchange=os.environ.get("GLU_BENCH_EXCHANGE", "auto"))

        def step_fn(k, v):
            last_out[0] = dsort(k, v, n)

        parallelism = (f"msd-split x{world}: top-8-bit histogram all-gather, balanced bucket->GPU prefix, "
                       f"{'fused partition + NVLink peer-store all-to-all' if dsort.exchange == 'p2p' else 'partition + NCCL all_to_all'}"
                       f", local onesweep sort")
    else:
        sorter = glu.RadixSort()
        sorter.prepare_internal_buffers(n)  # as the reference's benchmark does (test/radix_sort_tests.cpp:187)
        step_fn = lambda k, v: sorter(k, v, n)  # noqa: E731
        parallelism = "single GPU"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(warmup):
        step_fn(*inputs[i])
    barrier()
    glu.profile_enable(True)
    glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
    glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
    glu.profile_collect(glu.KERNEL_SORT_PARTITION)
    launches0 = glu.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    mark0 = sampler.mark()
    ev0.record()
    for i in range(steps):
        step_fn(*inputs[warmup + i])
    ev1.record()
    barrier()
    clocks = sampler.stop(mark0, sampler.mark())
    ms_total = ev0.elapsed_time(ev1)
    gpu_launches = glu.kernel_launch_count() - launches0
    sweep_ms, sweep_launches = glu.profile_collect(glu.KERNEL_SORT_ONESWEEP)
    hist_ms, hist_launches = glu.profile_collect(glu.KERNEL_SORT_HISTOGRAM)
    part_ms, part_launches = glu.profile_collect(glu.KERNEL_SORT_PARTITION)
    glu.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())

    # sanity: the last timed step really sorted its input
    k_sorted = last_out[0][0] if world > 1 else inputs[warmup + steps - 1][0]
    if world > 1:
        totals = torch.tensor([last_out[0][2]], dtype=torch.int64, device=dev)
        dist.all_reduce(totals)
        assert int(totals.item()) == world * n, "distributed sort lost or duplicated pairs"
    k64 = k_sorted[: 1 << 24].to(torch.int64) & 0xFFFFFFFF
    assert bool((k64[1:] >= k64[:-1]).all()), "bench output is not sorted"
    del k64

    value = world * n * steps / (ms_total * 1e-3) / 1e9
    # roofline of the dominant kernel (onesweep pass): algorithmic 16 B per pair per launch
    step_bytes = SORT_BYTES_PER_PAIR if world == 1 else SORT_BYTES_PER_PAIR + 4 + PASS_BYTES_PER_PAIR
    per_launch_ms = sweep_ms / max(1, sweep_launches)
    pairs_per_launch = n  # single GPU: every launch sweeps the whole array
    achieved = PASS_BYTES_PER_PAIR * pairs_per_launch / (per_launch_ms * 1e-3) / 1e9 if sweep_launches else None
    roofline = {"bound": "hbm", "kernel": "onesweep_kernel (one 8-bit digit pass)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": None,
                "peak_source": peak_src, "launches": sweep_launches, "ms_per_launch": per_launch_ms,
                "algorithmic_bytes_per_launch": PASS_BYTES_PER_PAIR * pairs_per_launch,
                "kernel_share_of_step": sweep_ms / ms_total,
                "histogram_ms_per_launch": hist_ms / max(1, hist_launches),
                "partition_exchange_ms_per_launch": (part_ms / part_launches) if part_launches else None,
                # per GPU and pair: the local sort's 68 B, plus at N > 1 the split-digit histogram (4 B) and the
                # partition pass (8 B read, 8 B written to local or peer memory)
                "whole_sort": {"bytes_per_pair": step_bytes,
                               "achieved_GB/s": step_bytes * n * steps / (ms_total * 1e-3) / 1e9,
                               "frac": step_bytes * n * steps / (ms_total * 1e-3) / 1e9 / peak}}
    traffic_file = os.path.join(ROOT, "profiles", "onesweep_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file))["dram_bytes_per_launch"]
        except Exception:
            pass

    line = {
        "metric": "radix_sort_u32_key_value_throughput", "value": value, "unit": "Gpairs/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"RadixSort of 2^{args.log2_pairs} uniform-random uint32 key/value pairs per GPU "
                               f"(BASELINE.json configs[2]), values = input index, one fresh unsorted input per step",
                   "pairs_per_gpu": n, "parallelism": parallelism,
                   "cache": "inputs (2 GiB per step) are larger than the 126 MB L2; no flush needed"},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(gpu_launches),
        "published_reference": {"value": 0.05345, "unit": "Gpairs/s", "hardware": "RTX 2060 SUPER (README.md:133)",
                                "note": "different hardware; not used for vs_baseline"},
    }

    if rank == 0 and world == 1 and not args.no_side_metrics:
        del inputs[1:]
        torch.cuda.empty_cache()
        # ---- e2e: the same sort through the host-buffer C-ABI, pinned host memory, copies inside the timed region.
        # One independent pinned (keys, values) input per step; the steps go through glu_host_sort_queue (depth 2):
        # the upload of step k+1 overlaps the sort and the download of step k, every step still pays its own
        # H2D + sort + D2H.  The synchronous single call (glu_radix_sort_u32kv_host) is timed beside it.
        e2e_steps = min(steps, 8)
        host = []
        for i in range(e2e_steps):
            hk, hk_ptr = pinned_u32(glu, n)
            hv, hv_ptr = pinned_u32(glu, n)
            hk[:] = np.random.default_rng(100 + i).integers(0, 1 << 32, size=n, dtype=np.uint32)
            hv[:] = np.arange(n, dtype=np.uint32)
            host.append((hk, hv, hk_ptr, hv_ptr))
        queue = glu.HostSortQueue(n, depth=2)
        wk, wk_ptr = pinned_u32(glu, 1 << 20)
        wv, wv_ptr = pinned_u32(glu, 1 << 20)
        wk[:] = np.arange(1 << 20, dtype=np.uint32)[::-1]
        wv[:] = 0
        queue.submit(wk, wv)  # warms the queue's streams
        queue.submit(wk, wv)
        queue.wait()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for hk, hv, _, _ in host:
            queue.submit(hk, hv, n)
        queue.wait()
        e2e_t = time.perf_counter() - t0
        queue.close()
        for hk, hv, _, _ in host:
            assert bool(np.all(hk[:-1][: 1 << 22] <= hk[1:][: 1 << 22])), "e2e output is not sorted"
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
        line["e2e"] = {"value": n * e2e_steps / e2e_t / 1e9, "unit": "Gpairs/s", "h2d_bytes_per_step": 8 * n,
                       "d2h_bytes_per_step": 8 * n, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_t / e2e_steps,
                       "api": "glu_host_sort_queue (depth 2; per step: pinned host arrays -> H2D -> sort -> D2H, "
                              "consecutive steps overlapped)",
                       "synchronous_call": {"value": n / sync_t / 1e9, "ms": 1e3 * sync_t,
                                            "api": "glu_radix_sort_u32kv_host"}}
        for _, _, p0, p1 in host:
            glu.lib.glu_free_host(p0)
            glu.lib.glu_free_host(p1)
        glu.lib.glu_free_host(wk_ptr)
        glu.lib.glu_free_host(wv_ptr)
        del host, hk, hv, wk, wv
        glu.profile_enable(True)
        line["side_metrics"] = side_metrics(glu, torch, dev, 1 << 28, peak)
        glu.profile_enable(False)
        line["cpu_baseline"] = cpu_baseline(args)
    elif world > 1 and not args.no_side_metrics:
        # ---- e2e at N GPUs: every rank uploads its shard from pinned host memory, the job sorts, every rank
        # downloads its slice of the result.  Like the single-GPU host queue, consecutive steps are software-pipelined:
        # the upload of step i+1 (copy stream, second device buffer) runs under the download of step i — PCIe is full
        # duplex — while every step still pays its own H2D + sort + D2H inside the timed region.  The download stays on
        # the sorting stream: the next step's partition pass writes into the peers' receive buffers, so it may only
        # start once everybody's slice has left them.  Wall clock between barriers, max over ranks.
        del inputs[1:]
        torch.cuda.empty_cache()
        e2e_steps = min(steps, 5)
        base_k, base_v = inputs[0][0].cpu(), inputs[0][1].cpu()
        try:  # 5.5 GiB of pinned host memory and 4 GiB of HBM per rank at 2^28 pairs
            hks = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]  # two unsorted shards, alternating
            hv = torch.empty(n, dtype=torch.int32).pin_memory()
            hv.copy_(base_v)
            for i, hk in enumerate(hks):
                hk.copy_(base_k ^ (0x9E3779B9 * (i + 1) & 0x7FFFFFFF))
            ok_ = torch.empty(dsort.capacity, dtype=torch.int32).pin_memory()
            ov_ = torch.empty(dsort.capacity, dtype=torch.int32).pin_memory()
            dks = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
            dvs = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
            allocated = 1
        except RuntimeError as e:  # e.g. the host cannot pin that much for every rank
            allocated = 0
            sys.stderr.write(f"[bench rank {rank}] e2e buffers: {e}\n")
        flag = torch.tensor([allocated], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks run the e2e leg or none does (it contains collectives)
        if int(flag.item()) == 0:
            line["e2e"] = None
            line["e2e_skipped"] = "a rank could not allocate the pinned host / device buffers of the e2e leg"
            if rank == 0:
                print(json.dumps(line), flush=True)
            dist.destroy_process_group()
            return
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        uploaded = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])  # the sort that read this device buffer two steps ago
                dks[i % 2].copy_(hks[i % 2], non_blocking=True)
                dvs[i % 2].copy_(hv, non_blocking=True)
                uploaded[i % 2].record(copy_stream)

        def run(first, count):
            upload(first)
            for i in range(first, first + count):
                main_stream.wait_event(uploaded[i % 2])
                if i + 1 < first + count:
                    upload(i + 1)
                sk, sv, m = dsort(dks[i % 2], dvs[i % 2], n)
                consumed[i % 2].record(main_stream)
                ok_[:m].copy_(sk, non_blocking=True)
                ov_[:m].copy_(sv, non_blocking=True)

        for ev in consumed:
            ev.record(main_stream)
        run(0, 1)  # warm-up step (pinned-copy paths, the copy stream)
        barrier()
        t0 = time.perf_counter()
        run(1, e2e_steps)
        barrier()
        e2e_t = time.perf_counter() - t0
        out64 = ok_[: 1 << 20].to(torch.int64) & 0xFFFFFFFF
        assert bool((out64[1:] >= out64[:-1]).all()), "e2e output is not sorted"
        t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
        line["e2e"] = {"value": world * n * e2e_steps / e2e_t / 1e9, "unit": "Gpairs/s",
                       "h2d_bytes_per_step": 8 * n * world, "d2h_bytes_per_step": 8 * n * world, "steps": e2e_steps,
                       "ms_per_step": 1e3 * e2e_t / e2e_steps,
                       "api": "DistributedRadixSort (per rank and step: pinned host shard -> H2D -> sort -> D2H of its "
                              "slice; the upload of step i+1 overlaps the download of step i)"}
    elif rank == 0:
        line["e2e"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
