"""GPU parity tests for glu::RadixSort's replacement — the cases of test/radix_sort_tests.cpp through the C ABI,
strengthened from the reference's keys-only is_sorted + permutation check to bit-equality of keys AND values
with std::stable_sort of the (key, value) pairs (oracle), value = input index."""
import numpy as np
import pytest

from conftest import fnv1a_u32, to_device, to_host

pytestmark = pytest.mark.gpu


def gpu_sort(glu, dev, keys: np.ndarray, vals: np.ndarray, num_steps=0, sorter=None):
    import torch

    dk, dv = to_device(keys, dev), to_device(vals, dev)
    (sorter or glu.RadixSort())(dk, dv, keys.size, num_steps)
    torch.cuda.synchronize()
    return to_host(dk, np.uint32), to_host(dv, np.uint32)


def check_against_oracle(oracle, keys, vals, gk, gv, num_steps=0):
    ek, ev = oracle.stable_sort_pairs(keys, vals, num_steps) if keys.size <= (1 << 22) else \
        oracle.lsd_sort_pairs(keys, vals, num_steps)
    np.testing.assert_array_equal(gk, ek)
    np.testing.assert_array_equal(gv, ev)


@pytest.mark.parametrize("n", [10, 128, 256, 512, 1024, 10993, 14978, 16243, 18985, 23857, 27865, 33363, 41298,
                               45821, 47487, 1048576])
def test_sort_reference_cases(glu, cuda_device, oracle, golden, n):
    # test/radix_sort_tests.cpp:54-110,136-158 (seed 1, "full range" = 31-bit keys) + BASELINE config 1 (2^20)
    keys = oracle.random_u32(1, n, 0, 0xFFFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals)
    # the reference's own assertions: permutation + sorted
    assert np.all(gk[:-1] <= gk[1:])
    np.testing.assert_array_equal(np.sort(gk), np.sort(keys))
    want = golden["sort_seed1"][str(n)]
    assert int(gk[0]) == want["min"] and int(gk[-1]) == want["max"]
    if n <= 50000:
        assert fnv1a_u32(gk) == want["keys_fnv1a"] and fnv1a_u32(gv) == want["vals_fnv1a"]
    check_against_oracle(oracle, keys, vals, gk, gv)


def test_sort_2048_heavy_duplicates(glu, cuda_device, oracle, golden):
    # test/radix_sort_tests.cpp:112-134 — keys in [0,10): stability decides the values
    keys = oracle.random_u32(1, 2048, 0, 10)
    vals = np.arange(2048, dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals)
    assert np.bincount(gk, minlength=10).tolist() == golden["sort_2048_digit_counts"]
    check_against_oracle(oracle, keys, vals, gk, gv)


@pytest.mark.parametrize("n", [2, 3, 31, 33, 2047, 2049, 6911, 6912, 6913, 100_000, (1 << 20) + 7, 3_000_001])
def test_sort_true_32bit_keys_ragged_sizes(glu, cuda_device, oracle, n):
    # bit 31 set (never exercised by the reference's generator), sizes around tile boundaries
    keys = oracle.mt19937_u32(1, n)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals)
    check_against_oracle(oracle, keys, vals, gk, gv)


def test_sort_matches_reference_algorithm(glu, cuda_device, oracle):
    # against the restated 8 x 4-bit GLSL algorithm itself, arbitrary (non-index) values
    keys = oracle.mt19937_u32(3, 30_000) & np.uint32(0x00FF00FF)
    vals = oracle.mt19937_u32(4, 30_000)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals)
    rk, rv, _ = oracle.radix_sort_glsl(keys, vals)
    np.testing.assert_array_equal(gk, rk)
    np.testing.assert_array_equal(gv, rv)


@pytest.mark.parametrize("num_steps", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_sort_num_steps(glu, cuda_device, oracle, num_steps):
    # glu/RadixSort.hpp:331 — only the low 4*num_steps bits take part; result always lands in the caller's
    # buffers (documented deviation: the reference leaves odd-num_steps results in its scratch)
    keys = oracle.mt19937_u32(7, 50_001)
    vals = np.arange(keys.size, dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals, num_steps)
    check_against_oracle(oracle, keys, vals, gk, gv, num_steps)
    rk, rv, _ = oracle.radix_sort_glsl(keys, vals, num_steps)
    np.testing.assert_array_equal(gk, rk)
    np.testing.assert_array_equal(gv, rv)


@pytest.mark.parametrize("kind", ["all_equal", "zero", "entropy16_low", "entropy16_high", "zipf", "sorted", "reversed",
                                  "two_values", "max_keys"])
def test_sort_skewed_inputs(glu, cuda_device, oracle, kind):
    # BASELINE config 5 flavours at a CPU-checkable size
    n = 1_500_000
    rng = np.random.default_rng(5)
    if kind == "all_equal":
        keys = np.full(n, 0xDEADBEEF, dtype=np.uint32)
    elif kind == "zero":
        keys = np.zeros(n, dtype=np.uint32)
    elif kind == "entropy16_low":
        keys = oracle.mt19937_u32(1, n) & np.uint32(0xFFFF)
    elif kind == "entropy16_high":
        keys = (oracle.mt19937_u32(1, n) & np.uint32(0xFFFF)) << np.uint32(16)
    elif kind == "zipf":
        keys = (rng.zipf(1.1, size=n) % (1 << 20)).astype(np.uint32)
    elif kind == "sorted":
        keys = np.sort(oracle.mt19937_u32(1, n))
    elif kind == "reversed":
        keys = np.sort(oracle.mt19937_u32(1, n))[::-1].copy()
    elif kind == "two_values":
        keys = (rng.integers(0, 2, size=n) * 0xFFFFFFFF).astype(np.uint32)
    else:
        keys = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
        keys[::3] = 0xFFFFFFFE
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys, vals)
    check_against_oracle(oracle, keys, vals, gk, gv)


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_sort_unaligned_buffers(glu, cuda_device, oracle, offset):
    import torch

    n = 40_000
    keys = oracle.mt19937_u32(8, n + offset)
    vals = np.arange(n + offset, dtype=np.uint32)
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    glu.RadixSort()(dk[offset:], dv[offset:], n)
    torch.cuda.synchronize()
    ek, ev = oracle.stable_sort_pairs(keys[offset:], vals[offset:])
    np.testing.assert_array_equal(to_host(dk, np.uint32)[offset:], ek)
    np.testing.assert_array_equal(to_host(dv, np.uint32)[offset:], ev)
    assert to_host(dk, np.uint32)[:offset].tolist() == keys[:offset].tolist()


def test_sort_object_reuse_and_prepare_internal_buffers(glu, cuda_device, oracle):
    # glu/RadixSort.hpp:237-271 — scratch is grow-only and reusable across calls of different sizes
    sorter = glu.RadixSort()
    sorter.prepare_internal_buffers(200_000)
    for n in (200_000, 1000, 150_000, 2):
        keys = oracle.mt19937_u32(n, n)
        vals = np.arange(n, dtype=np.uint32)
        gk, gv = gpu_sort(glu, cuda_device, keys, vals, sorter=sorter)
        check_against_oracle(oracle, keys, vals, gk, gv)


def test_sort_count_0_and_1_are_noops(glu, cuda_device):
    keys = np.array([5, 3], dtype=np.uint32)
    vals = np.array([0, 1], dtype=np.uint32)
    gk, gv = gpu_sort(glu, cuda_device, keys[:1], vals[:1])
    assert gk.tolist() == [5] and gv.tolist() == [0]


def test_sort_without_tma_path_matches(glu, cuda_device, oracle):
    # the cooperative ld/st staging path (taken for 16-byte-misaligned inputs and the last tile) on its own
    import subprocess, sys, os
    from conftest import ROOT
    code = (
        "import numpy as np, torch, __graft_entry__ as e, oracle\n"
        "glu = e.load_package()\n"
        "k = oracle.mt19937_u32(1, 500_000); v = np.arange(k.size, dtype=np.uint32)\n"
        "dk = torch.from_numpy(k.view(np.int32)).cuda(); dv = torch.from_numpy(v.view(np.int32)).cuda()\n"
        "glu.RadixSort()(dk, dv, k.size); torch.cuda.synchronize()\n"
        "ek, ev = oracle.stable_sort_pairs(k, v)\n"
        "assert np.array_equal(dk.cpu().numpy().view(np.uint32), ek) and np.array_equal(dv.cpu().numpy().view(np.uint32), ev)\n"
        "print('ok')\n")
    for env_extra in ({"GLU_SORT_TMA": "0"}, {"GLU_SORT_RANK": "0"}, {"GLU_SORT_CONFIG": "0"}, {"GLU_SORT_CONFIG": "4"}):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True)
        assert r.returncode == 0 and "ok" in r.stdout, (env_extra, r.stdout, r.stderr)


# GLU_SORT_SMALL_MAX=0: the tiny sizes go through the forced kernel form too (not through small_sort_kernel)
@pytest.mark.parametrize("env_extra", [{"GLU_SORT_CONFIG": "9", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "10", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "15", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "12", "GLU_SORT_TMA": "0", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "8", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "13", "GLU_SORT_CHAIN_ROWS": "104", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_CONFIG": "10", "GLU_SORT_RING_CTAS_PER_SM": "1", "GLU_SORT_SMALL_MAX": "0"},
                                       {"GLU_SORT_SMALL_MAX": "0"}, {}],
                         ids=lambda e: "-".join(f"{k[9:]}{v}" for k, v in e.items()))
def test_sort_both_kernel_forms(cuda_device, env_extra):
    """Every size class through BOTH forms of the digit pass, whatever the default selection is: the persistent ring
    kernel (GLU_SORT_CONFIG 9..18: tickets, two-deep key ring, early counts; 14..18 with the returning-atomic ranking
    loop) and the one-tile-per-CTA kernel (8).  One tile, a few tiles (more CTAs than tiles), many tiles, ragged
    ends, 16-byte-misaligned inputs (no bulk copies), heavy duplicates, num_steps, device-resident counts."""
    import os
    import subprocess
    import sys

    from conftest import ROOT
    code = r"""
import numpy as np, torch, __graft_entry__ as e, oracle
glu = e.load_package()
dev = torch.device('cuda', 0)
def up(a): return torch.from_numpy(a.view(np.int32).copy()).to(dev)
def check(k, v, num_steps=0, off=0):
    n = k.size - off
    dk, dv = up(k), up(v)
    glu.RadixSort()(dk[off:], dv[off:], n, num_steps); torch.cuda.synchronize()
    ek, ev = oracle.stable_sort_pairs(k[off:], v[off:], num_steps)
    assert np.array_equal(dk[off:].cpu().numpy().view(np.uint32), ek), (n, num_steps, off, 'keys')
    assert np.array_equal(dv[off:].cpu().numpy().view(np.uint32), ev), (n, num_steps, off, 'values')
for n in (2, 33, 5119, 5120, 5121, 7680, 15361, 41298, 100_003, 1_000_003, 3_000_001):
    k = oracle.mt19937_u32(n, n); v = np.arange(n, dtype=np.uint32)
    check(k, v)
n = 700_001
k = oracle.mt19937_u32(7, n); v = np.arange(n, dtype=np.uint32)
for off in (1, 2, 3):
    check(k, v, 0, off)
for steps in (1, 3, 5, 8):
    check(k, v, steps)
check(oracle.random_u32(1, n, 0, 10), v)                       # keys in [0, 10)
check(np.full(n, 0xDEADBEEF, dtype=np.uint32), v)              # one digit bin in every pass
check(oracle.mt19937_u32(3, n) & np.uint32(0xFFFF), v)         # 16-bit entropy
check(np.sort(k)[::-1].copy(), v)
# device-resident count (glu_radix_sort_u32kv_dyn): the scratch is sized for max_count, the kernels read the count
cap = 600_000
for m in (1, 2, 5121, 333_333, cap):
    kk = np.zeros(cap, np.uint32); kk[:m] = k[:m]
    dk, dv = up(kk), up(np.arange(cap, dtype=np.uint32))
    cnt = torch.tensor([m], dtype=torch.int32, device=dev)
    glu.RadixSort().sort_device_count(dk, dv, cnt, cap); torch.cuda.synchronize()
    ek, ev = oracle.stable_sort_pairs(k[:m], np.arange(m, dtype=np.uint32))
    assert np.array_equal(dk[:m].cpu().numpy().view(np.uint32), ek) and np.array_equal(dv[:m].cpu().numpy().view(np.uint32), ev), m
print('ok')
"""
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (env_extra, r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 255, 256, 257, 1000, 2047, 2048, 2049])
def test_sort_small_inputs_single_cta_path(glu, cuda_device, oracle, n):
    """Up to 2048 pairs are sorted by small_sort_kernel (one CTA, one launch); 2049 is the first size of the general
    path.  Full keys, num_steps, heavy duplicates, all-equal keys, unaligned buffers, and key-bit ranges through
    sort_ex — all against std::stable_sort of the pairs."""
    import torch

    vals = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(n)
    cases = [("uniform", oracle.mt19937_u32(n, n)), ("dups", oracle.random_u32(n, n, 0, 5)),
             ("equal", np.full(n, 0xFFFFFFFF, dtype=np.uint32)), ("ent16", oracle.mt19937_u32(n + 1, n) << np.uint32(16))]
    for name, keys in cases:
        for num_steps in (0, 1, 3, 5):
            dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
            glu.RadixSort()(dk, dv, n, num_steps)
            ek, ev = oracle.stable_sort_pairs(keys, vals, num_steps)
            np.testing.assert_array_equal(to_host(dk, np.uint32), ek, err_msg=f"{name} steps={num_steps} keys")
            np.testing.assert_array_equal(to_host(dv, np.uint32), ev, err_msg=f"{name} steps={num_steps} values")
    keys = oracle.mt19937_u32(7 * n, n + 3)
    vals3 = np.arange(n + 3, dtype=np.uint32)
    for off in (1, 3):   # 4-byte aligned only
        dk, dv = to_device(keys, cuda_device), to_device(vals3, cuda_device)
        glu.RadixSort()(dk[off:], dv[off:], n)
        ek, ev = oracle.stable_sort_pairs(keys[off:off + n], vals3[off:off + n])
        np.testing.assert_array_equal(to_host(dk, np.uint32)[off:off + n], ek)
        np.testing.assert_array_equal(to_host(dv, np.uint32)[off:off + n], ev)
        np.testing.assert_array_equal(to_host(dk, np.uint32)[:off], keys[:off])          # nothing outside [off, off + n)
        np.testing.assert_array_equal(to_host(dk, np.uint32)[off + n:], keys[off + n:])
    for begin_bit, end_bit in ((0, 32), (4, 20), (24, 32), (5, 6)):
        k = oracle.mt19937_u32(11 * n, n)
        dk, dv = to_device(k, cuda_device), to_device(vals, cuda_device)
        glu.RadixSort().sort_ex(dk, dv, n, begin_bit, end_bit)
        ek, ev = oracle.stable_sort_ex(k, vals, begin_bit, end_bit)
        np.testing.assert_array_equal(to_host(dk, np.uint32), ek, err_msg=f"bits [{begin_bit}, {end_bit}) keys")
        np.testing.assert_array_equal(to_host(dv, np.uint32), ev, err_msg=f"bits [{begin_bit}, {end_bit}) values")
    torch.cuda.synchronize()


def test_sort_full_size_2_28(glu, cuda_device, oracle):
    # BASELINE config 3: 2^28 uniform 32-bit keys, vals = index.  Size-independent properties on the device
    # (torch = plumbing): sorted, stable (equal keys => increasing source index), it IS the permutation it
    # claims (keys_in[val] == key_out), vals are a permutation (sum and xor of 0..n-1); plus exact CPU parity
    # of a 2^24 prefix slice sorted separately.
    import torch

    n = 1 << 28
    g = torch.Generator(device=cuda_device).manual_seed(1)
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    vals = torch.arange(n, dtype=torch.int32, device=cuda_device)
    dk, dv = keys.clone(), vals.clone()
    glu.RadixSort()(dk, dv, n)
    torch.cuda.synchronize()
    ok = True
    chunk = 1 << 26
    for i in range(0, n, chunk):
        j = min(n, i + chunk + 1)
        k64 = dk[i:j].to(torch.int64) & 0xFFFFFFFF
        v64 = dv[i:j].to(torch.int64)
        ok &= bool((k64[1:] >= k64[:-1]).all())
        ok &= bool(((k64[1:] > k64[:-1]) | (v64[1:] > v64[:-1])).all())  # stability
        ok &= bool(torch.equal(keys[dv[i:j].to(torch.int64)], dk[i:j]))
        del k64, v64
    assert ok
    assert int(dv.sum(dtype=torch.int64).item()) == n * (n - 1) // 2
    del dk, dv
    m = 1 << 24
    hk = to_host(keys[:m], np.uint32).copy()
    hv = np.arange(m, dtype=np.uint32)
    dk, dv = keys[:m].clone(), vals[:m].clone()
    glu.RadixSort()(dk, dv, m)
    ek, ev = oracle.lsd_sort_pairs(hk, hv)
    np.testing.assert_array_equal(to_host(dk, np.uint32), ek)
    np.testing.assert_array_equal(to_host(dv, np.uint32), ev)


@pytest.mark.parametrize("kind", ["zipf", "entropy16_low", "entropy16_high", "all_equal"])
def test_sort_skewed_full_size_2_28(glu, cuda_device, oracle, kind):
    """BASELINE configs[4] at the stated size: Zipf(1.1) over 2^20 distinct keys, 16-bit-entropy keys (low half and
    shifted into the top digit), all-equal keys; 2^28 pairs, values = input index.  Same on-device property checks as
    test_sort_full_size_2_28 (sorted, stable, it is the permutation it claims, values are a permutation) plus exact
    oracle parity of a separately sorted 2^24-pair slice of the same data."""
    import torch

    n = 1 << 28
    g = torch.Generator(device=cuda_device).manual_seed(11)
    if kind == "zipf":
        ranks = torch.arange(1, (1 << 20) + 1, dtype=torch.float64, device=cuda_device)
        cdf = torch.cumsum(ranks.pow(-1.1), 0)
        cdf /= cdf[-1].clone()
        keys = torch.empty(n, dtype=torch.int32, device=cuda_device)
        for i in range(0, n, 1 << 26):  # inverse-CDF sampling, chunked (float64 temporaries)
            u = torch.rand(1 << 26, dtype=torch.float64, device=cuda_device, generator=g)
            keys[i:i + (1 << 26)] = torch.searchsorted(cdf, u).to(torch.int32)
        del ranks, cdf, u
    elif kind == "entropy16_low":
        keys = torch.randint(0, 1 << 16, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    elif kind == "entropy16_high":
        keys = torch.randint(0, 1 << 16, (n,), dtype=torch.int32, device=cuda_device, generator=g) << 16
    else:
        keys = torch.full((n,), -559038737, dtype=torch.int32, device=cuda_device)  # 0xDEADBEEF
    vals = torch.arange(n, dtype=torch.int32, device=cuda_device)
    dk, dv = keys.clone(), vals.clone()
    glu.RadixSort()(dk, dv, n)
    torch.cuda.synchronize()
    ok = True
    chunk = 1 << 26
    for i in range(0, n, chunk):
        j = min(n, i + chunk + 1)
        k64 = dk[i:j].to(torch.int64) & 0xFFFFFFFF
        v64 = dv[i:j].to(torch.int64)
        ok &= bool((k64[1:] >= k64[:-1]).all())
        ok &= bool(((k64[1:] > k64[:-1]) | (v64[1:] > v64[:-1])).all())  # stability
        ok &= bool(torch.equal(keys[dv[i:j].to(torch.int64)], dk[i:j]))
        del k64, v64
    assert ok, kind
    assert int(dv.sum(dtype=torch.int64).item()) == n * (n - 1) // 2
    del dk, dv
    m = 1 << 24
    lo = n // 3  # a slice from the middle of the data
    hk = to_host(keys[lo:lo + m], np.uint32).copy()
    hv = np.arange(m, dtype=np.uint32)
    dk, dv = keys[lo:lo + m].clone(), vals[:m].clone()
    glu.RadixSort()(dk, dv, m)
    ek, ev = oracle.lsd_sort_pairs(hk, hv)
    np.testing.assert_array_equal(to_host(dk, np.uint32), ek)
    np.testing.assert_array_equal(to_host(dv, np.uint32), ev)


def test_sort_host_entry_point(glu, cuda_device, oracle):
    keys = oracle.mt19937_u32(12, 300_000)
    vals = np.arange(keys.size, dtype=np.uint32)
    hk, hv = keys.copy(), vals.copy()
    glu.radix_sort_u32kv_host(hk, hv)
    ek, ev = oracle.stable_sort_pairs(keys, vals)
    np.testing.assert_array_equal(hk, ek)
    np.testing.assert_array_equal(hv, ev)


def test_sort_host_queue_overlapped_jobs(glu, cuda_device, oracle):
    """glu_host_sort_queue_*: several host-buffer sorts in flight (depth 2, more jobs than slots, ragged sizes,
    one job limited by num_steps); every job must equal std::stable_sort of its own pairs."""
    sizes = [300_000, 1, 70_001, 1_000_003, 2, 555_555]
    q = glu.HostSortQueue(max(sizes), depth=2)
    jobs = []
    for j, n in enumerate(sizes):
        keys = oracle.mt19937_u32(40 + j, n)
        if j == 2:
            keys &= np.uint32(0xFFFF)
        vals = np.arange(n, dtype=np.uint32)
        hk, hv = keys.copy(), vals.copy()
        q.submit(hk, hv, n, 4 if j == 2 else 0)
        jobs.append((keys, vals, hk, hv))
    q.wait()
    for keys, vals, hk, hv in jobs:
        ek, ev = oracle.stable_sort_pairs(keys, vals)
        np.testing.assert_array_equal(hk, ek)
        np.testing.assert_array_equal(hv, ev)
    with pytest.raises(glu.GluError):
        q.submit(np.zeros(max(sizes) + 1, np.uint32), np.zeros(max(sizes) + 1, np.uint32))
    q.close()


def test_sort_beyond_2_30_pairs(glu, cuda_device):
    """2^30 + 2^27 pairs (BASELINE configs[3] puts 2^30 pairs on a GPU, and the multi-GPU receive buffers need head
    room above that): no CPU oracle at this size, so the size-independent properties — keys non-decreasing, equal keys
    keep their input order (values = index), and an order-independent checksum of the (key, value) pairs."""
    import torch

    if torch.cuda.get_device_properties(0).total_memory < 60 * (1 << 30):
        pytest.skip("needs ~30 GiB of device memory")
    n = (1 << 30) + (1 << 27)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    keys[: 1 << 20] = 7  # a long run of equal keys: stability must hold across many tiles
    vals = torch.arange(n, dtype=torch.int32, device=cuda_device)

    def checksum(k, v):
        total = 0
        for i in range(0, n, 1 << 28):  # chunked: the int64 temporaries are 2 GiB each
            kk = k[i:i + (1 << 28)].to(torch.int64) & 0xFFFFFFFF
            vv = v[i:i + (1 << 28)].to(torch.int64)
            total = (total + int(((kk * 0x9E3779B1 + vv * 0x85EBCA77) & 0xFFFFFFFFFFFF).sum().item())) & ((1 << 62) - 1)
        return total

    before = checksum(keys, vals)
    glu.RadixSort()(keys, vals, n)
    torch.cuda.synchronize()
    assert checksum(keys, vals) == before
    for i in range(0, n - 1, 1 << 28):
        j = min(n, i + (1 << 28) + 1)
        k = keys[i:j].to(torch.int64) & 0xFFFFFFFF
        v = vals[i:j]
        ok = (k[1:] > k[:-1]) | ((k[1:] == k[:-1]) & (v[1:] > v[:-1]))
        assert bool(ok.all()), f"order / stability violated in chunk starting at {i}"
