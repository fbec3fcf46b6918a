"""GPU parity tests for glu_radix_sort_u32_ex (key-only sort, key-bit ranges, descending order).
The reference has none of these (README.md:88-89 lists the mandatory value buffer as a limitation), so the oracle is
std::stable_sort with the matching comparator (oracle.stable_sort_ex) — parity unpinned by the reference, bit-exact
against the oracle.  Everything goes through the C ABI."""
import numpy as np
import pytest

from conftest import to_device, to_host

pytestmark = pytest.mark.gpu

# sizes on both sides of the three tile shapes (2048 / 4096 / 7680 pairs) and of their size thresholds (2^18, 2^21)
SIZES = [2, 33, 2047, 2049, 4097, 7679, 7681, 100_003, (1 << 18) + 5, (1 << 21) + 7681 + 3]


def gpu_sort_ex(glu, dev, keys, vals, begin_bit=0, end_bit=32, descending=False, sorter=None):
    import torch

    dk = to_device(keys, dev)
    dv = to_device(vals, dev) if vals is not None else None
    (sorter or glu.RadixSort()).sort_ex(dk, dv, keys.size, begin_bit, end_bit, descending)
    torch.cuda.synchronize()
    return to_host(dk, np.uint32), (to_host(dv, np.uint32) if dv is not None else None)


def check(oracle, keys, vals, gk, gv, begin_bit=0, end_bit=32, descending=False):
    ek, ev = oracle.stable_sort_ex(keys, vals, begin_bit, end_bit, descending)
    np.testing.assert_array_equal(gk, ek)
    if vals is not None:
        np.testing.assert_array_equal(gv, ev)


@pytest.mark.parametrize("n", SIZES)
def test_sort_ex_defaults_equal_the_plain_sort(glu, cuda_device, oracle, n):
    keys = oracle.mt19937_u32(3, n)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals)
    ek, ev = oracle.stable_sort_pairs(keys, vals)
    np.testing.assert_array_equal(gk, ek)
    np.testing.assert_array_equal(gv, ev)


@pytest.mark.parametrize("n", SIZES)
def test_sort_ex_keys_only(glu, cuda_device, oracle, n):
    keys = oracle.mt19937_u32(4, n)
    gk, _ = gpu_sort_ex(glu, cuda_device, keys, None)
    np.testing.assert_array_equal(gk, np.sort(keys))


@pytest.mark.parametrize("n", SIZES)
def test_sort_ex_descending_pairs_is_stable(glu, cuda_device, oracle, n):
    # few distinct keys: stability decides the values
    keys = (oracle.mt19937_u32(5, n) % np.uint32(97)) * np.uint32(44_000_003)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, descending=True)
    assert np.all(gk[:-1] >= gk[1:])
    check(oracle, keys, vals, gk, gv, descending=True)


@pytest.mark.parametrize("n", [2049, 100_003, (1 << 21) + 11])
def test_sort_ex_descending_keys_only(glu, cuda_device, oracle, n):
    keys = oracle.mt19937_u32(6, n)
    gk, _ = gpu_sort_ex(glu, cuda_device, keys, None, descending=True)
    np.testing.assert_array_equal(gk, np.sort(keys)[::-1])


@pytest.mark.parametrize("begin_bit,end_bit", [(0, 8), (0, 12), (8, 16), (8, 32), (5, 17), (24, 32), (31, 32), (3, 4),
                                               (0, 31), (16, 25)])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_ex_bit_ranges(glu, cuda_device, oracle, begin_bit, end_bit, descending):
    n = 300_007
    keys = oracle.mt19937_u32(7, n)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, begin_bit, end_bit, descending)
    check(oracle, keys, vals, gk, gv, begin_bit, end_bit, descending)


def test_sort_ex_num_steps_equivalence(glu, cuda_device, oracle):
    # glu::RadixSort's num_steps (glu/RadixSort.hpp:331) == bit range [0, 4 * num_steps)
    n = 50_001
    keys = oracle.mt19937_u32(8, n)
    vals = np.arange(n, dtype=np.uint32)
    for num_steps in (1, 3, 5, 7):
        gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, 0, 4 * num_steps)
        ek, ev = oracle.stable_sort_pairs(keys, vals, num_steps)
        np.testing.assert_array_equal(gk, ek)
        np.testing.assert_array_equal(gv, ev)


def test_sort_ex_empty_bit_range_and_tiny_counts_are_noops(glu, cuda_device, oracle):
    keys = oracle.mt19937_u32(9, 1000)
    vals = np.arange(1000, dtype=np.uint32)
    gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, 13, 13)
    np.testing.assert_array_equal(gk, keys)
    np.testing.assert_array_equal(gv, vals)
    gk, _ = gpu_sort_ex(glu, cuda_device, keys[:1], None)
    np.testing.assert_array_equal(gk, keys[:1])


@pytest.mark.parametrize("kind", ["all_equal", "entropy16_high", "sorted", "reversed"])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_ex_skewed_inputs(glu, cuda_device, oracle, kind, descending):
    n = 200_003
    r = oracle.mt19937_u32(10, n)
    keys = {"all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
            "entropy16_high": (r & np.uint32(0xFFFF)) << np.uint32(16),
            "sorted": np.sort(r),
            "reversed": np.sort(r)[::-1].copy()}[kind]
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, descending=descending)
    check(oracle, keys, vals, gk, gv, descending=descending)
    gk, _ = gpu_sort_ex(glu, cuda_device, keys, None, descending=descending)
    np.testing.assert_array_equal(gk, np.sort(keys)[::-1] if descending else np.sort(keys))


@pytest.mark.parametrize("offset", [1, 3])
def test_sort_ex_unaligned_buffers(glu, cuda_device, oracle, offset):
    # not 16-byte aligned: the cooperative staging path instead of the TMA bulk copies
    import torch

    n = 30_011
    keys = oracle.mt19937_u32(11, n)
    vals = np.arange(n, dtype=np.uint32)
    bk = torch.zeros(n + 8, dtype=torch.int32, device=cuda_device)
    bv = torch.zeros(n + 8, dtype=torch.int32, device=cuda_device)
    dk, dv = bk[offset:offset + n], bv[offset:offset + n]
    dk.copy_(to_device(keys, cuda_device))
    dv.copy_(to_device(vals, cuda_device))
    s = glu.RadixSort()
    s.sort_ex(dk, dv, n, 0, 32, True)
    torch.cuda.synchronize()
    check(oracle, keys, vals, to_host(dk, np.uint32), to_host(dv, np.uint32), descending=True)
    dk.copy_(to_device(keys, cuda_device))
    s.sort_ex(dk, None, n)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(to_host(dk, np.uint32), np.sort(keys))


def test_sort_ex_object_reuse_across_flavours(glu, cuda_device, oracle):
    # one object, growing and shrinking scratch needs, every flavour after the other
    s = glu.RadixSort()
    for n, with_vals, desc in [(5000, True, False), (400_000, False, True), (70_000, True, True), (2_200_000, False, False),
                               (9000, True, False)]:
        keys = oracle.mt19937_u32(12 + n % 7, n)
        vals = np.arange(n, dtype=np.uint32) if with_vals else None
        gk, gv = gpu_sort_ex(glu, cuda_device, keys, vals, descending=desc, sorter=s)
        check(oracle, keys, vals, gk, gv, descending=desc)
    # and the plain entry point on the same object afterwards
    keys = oracle.mt19937_u32(1, 123_457)
    vals = np.arange(keys.size, dtype=np.uint32)
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    s(dk, dv, keys.size)
    ek, ev = oracle.stable_sort_pairs(keys, vals)
    np.testing.assert_array_equal(to_host(dk, np.uint32), ek)
    np.testing.assert_array_equal(to_host(dv, np.uint32), ev)


def test_sort_ex_keys_only_2_28(glu, cuda_device):
    # full size through size-independent properties: sortedness (descending) and a permutation checksum
    import torch

    n = 1 << 28
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(28)
    k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=cuda_device, generator=gen)
    k64 = k.to(torch.int64) & 0xFFFFFFFF
    s0, x0 = int(k64.sum()), int(torch.bitwise_xor(k64[: n // 2], k64[n // 2:]).sum())
    del k64
    glu.RadixSort().sort_ex(k, None, n, 0, 32, True)
    torch.cuda.synchronize()
    u = k.to(torch.int64) & 0xFFFFFFFF
    assert bool((u[:-1] >= u[1:]).all()), "not sorted descending"
    assert int(u.sum()) == s0, "key multiset changed"
    hist0 = torch.bincount((u >> 24).to(torch.int64), minlength=256)
    assert int(hist0.sum()) == n
