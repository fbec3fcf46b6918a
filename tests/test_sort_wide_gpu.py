"""GPU parity tests for glu_radix_sort_wide (64-bit keys, payloads wider than 32 bits).  Beyond the
reference (uint32 keys + mandatory uint32 value, README.md:88-89): parity unpinned by it; the oracle is a stable sort of
the keys with the value rows carried along (oracle.stable_sort_wide).  Bit-exact, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [2, 33, 2049, 100_003, (1 << 21) + 7681 + 3]


def _to_dev(a, dev):
    import torch

    signed = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}.get(a.dtype)
    return torch.from_numpy((a.view(signed) if signed else a).copy()).to(dev)


def _to_host(t, dtype):
    return t.detach().cpu().numpy().view(dtype)


def make_keys(oracle, seed, n, key_bytes, few_distinct=False):
    lo = oracle.mt19937_u32(seed, n)
    if key_bytes == 4:
        return lo % np.uint32(1000) if few_distinct else lo
    hi = oracle.mt19937_u32(seed + 100, n)
    if few_distinct:  # many equal high words AND many fully equal keys: both sorts' stability is needed
        hi, lo = hi % np.uint32(7), lo % np.uint32(5)
    return (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)


def make_vals(n, value_bytes):
    if value_bytes == 0:
        return None
    idx = np.arange(n, dtype=np.uint32)
    if value_bytes == 4:
        return idx
    if value_bytes == 8:
        return idx.astype(np.uint64) * np.uint64(0x1_0000_0001) + np.uint64(7)
    return np.stack([idx, ~idx, idx * np.uint32(3), idx ^ np.uint32(0xABCD1234)], axis=1).copy()  # 16-byte rows


def run(glu, dev, oracle, keys, vals, key_bytes, value_bytes, descending, sorter=None):
    import torch

    dk = _to_dev(keys, dev)
    dv = _to_dev(vals, dev) if vals is not None else None
    (sorter or glu.RadixSort()).sort_wide(dk, dv, keys.size, key_bytes, value_bytes, descending)
    torch.cuda.synchronize()
    ek, ev = oracle.stable_sort_wide(keys, vals, descending)
    np.testing.assert_array_equal(_to_host(dk, keys.dtype), ek)
    if vals is not None:
        np.testing.assert_array_equal(_to_host(dv, vals.dtype).reshape(vals.shape), ev)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("value_bytes", [0, 4, 8, 16])
def test_sort_wide_u64_keys(glu, cuda_device, oracle, n, value_bytes):
    keys = make_keys(oracle, 31, n, 8)
    run(glu, cuda_device, oracle, keys, make_vals(n, value_bytes), 8, value_bytes, False)


@pytest.mark.parametrize("n", [2049, 300_007])
@pytest.mark.parametrize("value_bytes", [0, 4, 16])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_wide_u64_keys_duplicates_are_stable(glu, cuda_device, oracle, n, value_bytes, descending):
    keys = make_keys(oracle, 32, n, 8, few_distinct=True)
    run(glu, cuda_device, oracle, keys, make_vals(n, value_bytes), 8, value_bytes, descending)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("value_bytes", [8, 16])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_wide_u32_keys_wide_values(glu, cuda_device, oracle, n, value_bytes, descending):
    keys = make_keys(oracle, 33, n, 4, few_distinct=(n % 2 == 1))
    run(glu, cuda_device, oracle, keys, make_vals(n, value_bytes), 4, value_bytes, descending)


@pytest.mark.parametrize("value_bytes", [0, 4])
def test_sort_wide_narrow_cases_forward_to_the_u32_sort(glu, cuda_device, oracle, value_bytes):
    keys = make_keys(oracle, 34, 70_001, 4)
    run(glu, cuda_device, oracle, keys, make_vals(keys.size, value_bytes), 4, value_bytes, True)


def test_sort_wide_special_key_values_and_object_reuse(glu, cuda_device, oracle):
    s = glu.RadixSort()
    n = 50_000
    keys = make_keys(oracle, 35, n, 8)
    keys[:8] = np.array([0, 1, 0xFFFFFFFF, 0x1_0000_0000, 0xFFFFFFFF_00000000, 0xFFFFFFFF_FFFFFFFF, 1 << 63, (1 << 63) - 1],
                        dtype=np.uint64)
    for desc in (False, True):
        run(glu, cuda_device, oracle, keys, make_vals(n, 8), 8, 8, desc, sorter=s)
    run(glu, cuda_device, oracle, make_keys(oracle, 36, 3000, 4), make_vals(3000, 16), 4, 16, False, sorter=s)
    run(glu, cuda_device, oracle, keys[:1], make_vals(1, 4), 8, 4, False, sorter=s)  # count == 1: no-op


def test_sort_wide_u64_2_26_properties(glu, cuda_device):
    # 2^26 pairs through size-independent properties: sortedness of the 64-bit keys, (key, value) pairing intact
    import torch

    n = 1 << 26
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(26)
    k = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device=cuda_device, generator=gen)
    k = k * 2 + torch.randint(0, 2, (n,), dtype=torch.int64, device=cuda_device, generator=gen)  # all 64 bits vary
    v = (k ^ 0x5DEECE66D).clone()  # 8-byte value tied to its key
    glu.RadixSort().sort_wide(k, v, n, 8, 8, False)
    torch.cuda.synchronize()
    assert bool(((v ^ 0x5DEECE66D) == k).all()), "values no longer belong to their keys"
    # unsigned order: compare (high bit, rest) lexicographically == signed compare after flipping the sign bit
    u = k ^ (-(1 << 63))
    assert bool((u[1:] >= u[:-1]).all()), "64-bit keys are not sorted"
