"""GPU parity tests for glu::BlellochScan's replacement — the cases of test/blelloch_scan_tests.cpp through the
C ABI against std::exclusive_scan (oracle), plus what the reference cannot do (non powers of two)."""
import numpy as np
import pytest

from conftest import fnv1a_u32, to_device, to_host

pytestmark = pytest.mark.gpu


def gpu_scan(glu, dev, data: np.ndarray, data_type, count=None, parts=1):
    import torch

    buf = to_device(data.reshape(-1), dev)
    if count is None:
        count = data.shape[0]
    glu.BlellochScan(data_type)(buf, count, parts)
    torch.cuda.synchronize()
    return to_host(buf, data.dtype).reshape(data.shape)


def test_scan_simple(glu, cuda_device):
    # test/blelloch_scan_tests.cpp:12-26
    data = np.arange(1, 9, dtype=np.uint32)
    assert gpu_scan(glu, cuda_device, data, glu.DataType_Uint).tolist() == [0, 1, 3, 6, 10, 15, 21, 28]


@pytest.mark.parametrize("n", [1 << k for k in range(10, 21)])
def test_scan_multiple_sizes(glu, cuda_device, oracle, golden, n):
    # test/blelloch_scan_tests.cpp:28-46 — seed 123, [0,100), powers of two 2^10..2^20
    data = oracle.random_u32(123, n, 0, 100)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Uint)
    np.testing.assert_array_equal(got, oracle.exclusive_scan(data))
    assert int(got[-1]) == golden["scan_seed123_0_100"][str(n)]["last"]
    if n <= 65536:
        assert fnv1a_u32(got) == golden["scan_seed123_0_100"][str(n)]["fnv1a"]


@pytest.mark.parametrize("parts", [1, 32, 100, 1000])
def test_scan_multiple_partitions(glu, cuda_device, oracle, parts):
    # test/blelloch_scan_tests.cpp:48-82 — 1024 elements per partition
    data = oracle.random_u32(123, 1024 * parts, 0, 100)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Uint, 1024, parts)
    np.testing.assert_array_equal(got, oracle.exclusive_scan(data, 1024, parts))


@pytest.mark.parametrize("count,parts", [(1, 1), (1, 77), (3, 5), (500, 16), (1025, 16), (4095, 3), (4097, 3),
                                         (70_001, 7), (262144, 16), (1 << 20, 2), (999_999, 1), (12_345_679, 1)])
def test_scan_arbitrary_counts_and_partitions(glu, cuda_device, oracle, count, parts):
    # the reference rejects non powers of two (glu/BlellochScan.hpp:134); full-range values wrap mod 2^32
    data = oracle.mt19937_u32(17, count * parts)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Uint, count, parts)
    np.testing.assert_array_equal(got, oracle.exclusive_scan(data, count, parts))


def test_scan_matches_shader_restatement(glu, cuda_device, oracle):
    # same result as the reference's real up-sweep / down-sweep algorithm, incl. its sort use (16 partitions)
    data = oracle.mt19937_u32(2, 2048 * 16)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Uint, 2048, 16)
    np.testing.assert_array_equal(got, oracle.blelloch_scan_glsl(data, 2048, 16))


def test_scan_int(glu, cuda_device, oracle):
    data = oracle.mt19937_u32(4, 300_000).view(np.int32)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Int)
    np.testing.assert_array_equal(got, oracle.exclusive_scan(data))


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_scan_unaligned_buffer(glu, cuda_device, oracle, offset):
    import torch

    data = oracle.mt19937_u32(6, 50_000 + offset)
    buf = to_device(data, cuda_device)
    view = buf[offset:]
    glu.BlellochScan(glu.DataType_Uint)(view, view.numel())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(to_host(view, np.uint32), oracle.exclusive_scan(data[offset:]))
    assert to_host(buf, np.uint32)[:offset].tolist() == data[:offset].tolist()


@pytest.mark.parametrize("count,parts", [(1000, 1), (100_003, 1), (4096, 9), (1 << 21, 1)])
def test_scan_float(glu, cuda_device, oracle, count, parts):
    rng = np.random.default_rng(8)
    data = rng.uniform(-1.0, 1.0, size=count * parts).astype(np.float32)
    got = gpu_scan(glu, cuda_device, data, glu.DataType_Float, count, parts)
    want = oracle.exclusive_scan(data, count, parts)
    # tolerance: 1e-6 relative to the running sum of |x| (float32 scan, any association order)
    scale = np.cumsum(np.abs(data.astype(np.float64)).reshape(parts, count), axis=1).reshape(-1)
    assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= 1e-6 * scale + 1e-7)
    assert not got.reshape(parts, count)[:, 0].any()


@pytest.mark.parametrize("dt_name,np_dtype,ncomp", [("Double", np.float64, 1), ("Vec2", np.float32, 2),
                                                    ("Vec4", np.float32, 4), ("DVec2", np.float64, 2),
                                                    ("DVec4", np.float64, 4)])
@pytest.mark.parametrize("count,parts", [(10, 1), (5000, 3), (200_001, 1)])
def test_scan_wide_floating_types(glu, cuda_device, oracle, dt_name, np_dtype, ncomp, count, parts):
    rng = np.random.default_rng(9)
    data = rng.uniform(-1.0, 1.0, size=(count * parts, ncomp)).astype(np_dtype)
    got = gpu_scan(glu, cuda_device, data, getattr(glu, "DataType_" + dt_name), count, parts)
    rel = 1e-6 if np_dtype == np.float32 else 1e-13
    for c in range(ncomp):
        col = np.ascontiguousarray(data[:, c])
        want = oracle.exclusive_scan(col, count, parts)
        scale = np.cumsum(np.abs(col.astype(np.float64)).reshape(parts, count), axis=1).reshape(-1)
        assert np.all(np.abs(got[:, c].astype(np.float64) - want.astype(np.float64)) <= rel * scale + 1e-7)


@pytest.mark.parametrize("dt_name,np_dtype,ncomp", [("UVec2", np.uint32, 2), ("UVec4", np.uint32, 4),
                                                    ("IVec2", np.int32, 2), ("IVec4", np.int32, 4)])
def test_scan_integer_vectors(glu, cuda_device, oracle, dt_name, np_dtype, ncomp):
    count, parts = 33_333, 4
    data = oracle.mt19937_u32(31, count * parts * ncomp).reshape(count * parts, ncomp).view(np_dtype)
    got = gpu_scan(glu, cuda_device, data, getattr(glu, "DataType_" + dt_name), count, parts)
    for c in range(ncomp):
        col = np.ascontiguousarray(data[:, c])
        np.testing.assert_array_equal(got[:, c], oracle.exclusive_scan(col, count, parts))


def test_scan_full_size_2_28(glu, cuda_device, oracle):
    # BASELINE config 2: BlellochScan(Uint) over 2^28 uint32. Device-side check against torch.cumsum
    # (int64, plumbing only) + CPU-oracle check of the leading 2^24 elements.
    import torch

    n = 1 << 28
    g = torch.Generator(device=cuda_device).manual_seed(123)
    t = torch.randint(0, 100, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    head = to_host(t[: 1 << 24], np.uint32).copy()
    buf = t.clone()
    glu.BlellochScan(glu.DataType_Uint)(buf, n)
    np.testing.assert_array_equal(to_host(buf[: 1 << 24], np.uint32), oracle.exclusive_scan(head))
    ok = True
    chunk = 1 << 26
    carry = 0
    for i in range(0, n, chunk):
        inc = torch.cumsum(t[i:i + chunk], 0, dtype=torch.int64)
        exc = inc - t[i:i + chunk] + carry
        ok &= bool(torch.equal(exc & 0xFFFFFFFF, buf[i:i + chunk].to(torch.int64) & 0xFFFFFFFF))
        carry += int(inc[-1].item())
        del inc, exc
    assert ok


def test_scan_host_entry_point(glu, cuda_device, oracle):
    data = oracle.mt19937_u32(10, 777_777)
    buf = data.copy()
    glu.scan_exclusive_host(buf, buf.size, 1, glu.DataType_Uint)
    np.testing.assert_array_equal(buf, oracle.exclusive_scan(data))
