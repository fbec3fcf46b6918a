"""CPU tests of bench.py's contract (no GPU): the reference arm runs here, prints ONE JSON line with the keys the driver
reads, names the same workload as the B200 arm, and under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample-log2", "16", *args], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "radix_sort_u32_key_value_throughput" and d["unit"] == "Gpairs/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "glsl_llvmpipe" in d and d["gpu_launches"] == 0
    # the workload named is the B200 arm's (same function builds both `config` objects)
    sys.path.insert(0, ROOT)
    import bench

    class A:
        log2_pairs = 28
    assert d["config"] == bench.workload_config(A)
    assert set(d["config"]) == {"workload", "pairs_per_gpu", "parallelism", "cache"}


def test_reference_arm_ignores_torchrun_thread_cap_and_other_ranks_stay_silent():
    d = json.loads(_run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")[0])
    assert d["n_gpus"] == 2
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))   # not torchrun's OMP_NUM_THREADS=1
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2") == []     # only rank 0 prints
