"""GPU parity tests for glu::Reduce's replacement — same cases as test/reduce_tests.cpp, through the C ABI,
checked against the oracle (std::accumulate / min / max) and the reference's literal known answers."""
import numpy as np
import pytest

from conftest import to_device, to_host
from test_oracle import K_DATA_100

pytestmark = pytest.mark.gpu


def gpu_reduce(glu, dev, data: np.ndarray, data_type, op, count=None):
    import torch

    buf = to_device(data.reshape(-1), dev)
    if count is None:
        count = data.shape[0]
    glu.Reduce(data_type, op)(buf, count)
    torch.cuda.synchronize()
    return to_host(buf, data.dtype)


def test_reduce_simple_uint(glu, cuda_device):
    # test/reduce_tests.cpp:14-52
    assert gpu_reduce(glu, cuda_device, K_DATA_100, glu.DataType_Uint, glu.ReduceOperator_Sum)[0] == 4951
    assert gpu_reduce(glu, cuda_device, K_DATA_100, glu.DataType_Uint, glu.ReduceOperator_Mul, 5)[0] == 319200
    assert gpu_reduce(glu, cuda_device, K_DATA_100, glu.DataType_Uint, glu.ReduceOperator_Min)[0] == 1
    assert gpu_reduce(glu, cuda_device, K_DATA_100, glu.DataType_Uint, glu.ReduceOperator_Max)[0] == 99


def test_reduce_all(glu, cuda_device):
    # test/reduce_tests.cpp:54-145 (tolerance +-0.1 absolute for floating types, as in the reference)
    dev = cuda_device
    S = glu.ReduceOperator_Sum
    u = np.array([1, 11, 80, 73, 48, 40, 89, 36, 70, 57], dtype=np.uint32)
    assert gpu_reduce(glu, dev, u, glu.DataType_Uint, S)[0] == 505
    f = np.array([42.138, 18.228, -19.127, 86.564, 11.904, 48.538, 30.606, 11.338, -32.699, -29.587], dtype=np.float32)
    assert abs(gpu_reduce(glu, dev, f, glu.DataType_Float, S)[0] - 167.9) < 0.1
    d = np.array([-6.20, -56.02, 49.42, 52.38, -23.81, -29.72, 95.46, 77.37, -85.00, 81.74], dtype=np.float64)
    assert abs(gpu_reduce(glu, dev, d, glu.DataType_Double, S)[0] - 155.6) < 0.1
    v2 = np.array([[-77.08, 19.54], [98.89, -16.09], [10.53, 91.17], [43.06, -94.18], [-19.18, 0.86],
                   [-49.99, -92.53], [-4.68, 42.34], [2.79, -4.26], [-17.49, 43.99], [79.45, -14.58]], dtype=np.float32)
    np.testing.assert_allclose(gpu_reduce(glu, dev, v2, glu.DataType_Vec2, S)[:2], [66.29, -23.75], atol=0.1)
    v4 = np.array([[-17.04, 1.79, 82.67, 39.72], [52.66, 24.75, -19.05, 91.92], [19.15, 44.93, -52.13, 18.85],
                   [-84.25, 69.53, -11.43, 33.17], [19.46, -14.30, -15.20, -63.83], [-20.51, -56.75, -2.70, 82.66],
                   [3.86, 55.48, -12.37, -11.02], [-30.62, -67.54, -29.89, -77.30], [-21.55, 50.46, 39.34, 81.08],
                   [-56.40, 84.61, 90.26, 13.35]], dtype=np.float32)
    np.testing.assert_allclose(gpu_reduce(glu, dev, v4, glu.DataType_Vec4, S)[:4], [-135.24, 192.97, 69.49, 208.59],
                               atol=0.1)
    i2 = np.array([[-38, -88], [57, -34], [61, 60], [-90, 73], [-23, -17], [34, -79], [-80, 53], [24, -23],
                   [-88, 69], [-83, -67]], dtype=np.int32)
    assert gpu_reduce(glu, dev, i2, glu.DataType_IVec2, S)[:2].tolist() == [-226, -53]
    i4 = np.array([[-95, 99, -30, 2], [-69, 33, 78, 20], [33, -43, -38, -26], [69, -67, -17, -57],
                   [18, -23, -2, -53], [88, -96, 40, -48], [-93, -47, -91, 59], [-89, 82, 10, 94],
                   [-15, 7, 41, 14], [63, 53, -40, 53]], dtype=np.int32)
    assert gpu_reduce(glu, dev, i4, glu.DataType_IVec4, S)[:4].tolist() == [-90, -2, -49, 58]


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072,
                               1, 31, 93, 201, 693, 2087, 7358, 88289, 345897, 6094798, 5238082, 10043898])
def test_reduce_seeded_sizes(glu, cuda_device, oracle, golden, n):
    # test/reduce_tests.cpp:147-183 (subgroup fitting and non-fitting sizes), seed 1, values [0,100)
    data = oracle.random_u32(1, n, 0, 100)
    want = golden["reduce_seed1_0_100"][str(n)]
    got = gpu_reduce(glu, cuda_device, data, glu.DataType_Uint, glu.ReduceOperator_Sum)
    assert int(got[0]) == want["sum"] == oracle.reduce(data, oracle.OP_SUM)
    if n > 1:
        np.testing.assert_array_equal(got[1:], data[1:])  # elements past 0 are left untouched
    assert int(gpu_reduce(glu, cuda_device, data, glu.DataType_Uint, glu.ReduceOperator_Min)[0]) == want["min"]
    assert int(gpu_reduce(glu, cuda_device, data, glu.DataType_Uint, glu.ReduceOperator_Max)[0]) == want["max"]


@pytest.mark.parametrize("op", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 1000, 4097, 1 << 20, (1 << 22) + 3])
def test_reduce_uint_int_all_ops_wraparound(glu, cuda_device, oracle, n, op):
    data = oracle.mt19937_u32(11, n)  # full 32-bit range: sums and products wrap mod 2^32
    if op == 1:
        data |= np.uint32(1)  # odd factors keep the product from collapsing to 0
    got = gpu_reduce(glu, cuda_device, data, glu.DataType_Uint, op)
    assert int(got[0]) == oracle.reduce(data, op)
    idata = data.view(np.int32)
    goti = gpu_reduce(glu, cuda_device, idata, glu.DataType_Int, op)
    assert int(goti[0]) == oracle.reduce(idata, op)


@pytest.mark.parametrize("offset", [1, 2, 3])
def test_reduce_unaligned_buffer(glu, cuda_device, oracle, offset):
    import torch

    data = oracle.mt19937_u32(5, 100_000 + offset)
    buf = to_device(data, cuda_device)
    view = buf[offset:]
    glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)(view, view.numel())
    torch.cuda.synchronize()
    assert int(to_host(view, np.uint32)[0]) == oracle.reduce(data[offset:], oracle.OP_SUM)


@pytest.mark.parametrize("dt_name,np_dtype,ncomp", [("Float", np.float32, 1), ("Double", np.float64, 1),
                                                    ("Vec2", np.float32, 2), ("Vec4", np.float32, 4),
                                                    ("DVec2", np.float64, 2), ("DVec4", np.float64, 4)])
@pytest.mark.parametrize("n", [7, 1000, 300_001])
def test_reduce_floating_types(glu, cuda_device, oracle, dt_name, np_dtype, ncomp, n):
    rng = np.random.default_rng(42)
    data = rng.uniform(-1.0, 1.0, size=(n, ncomp)).astype(np_dtype)
    dt = getattr(glu, "DataType_" + dt_name)
    # Min / Max are exact; Sum within 1e-5 * sum|x| (f32) / 1e-12 * sum|x| (f64) of the long-accumulated CPU sum
    for op, tol_scale in ((2, 0.0), (3, 0.0), (0, 1e-5 if np_dtype == np.float32 else 1e-12)):
        got = gpu_reduce(glu, cuda_device, data, dt, op)[:ncomp]
        for c in range(ncomp):
            col = np.ascontiguousarray(data[:, c])
            want = oracle.reduce(col, op)
            tol = tol_scale * float(np.abs(col.astype(np.float64)).sum())
            assert abs(float(got[c]) - want) <= tol, (dt_name, op, c, got[c], want)
    # Mul on a short, well-conditioned vector
    small = rng.uniform(0.9, 1.1, size=(50, ncomp)).astype(np_dtype)
    got = gpu_reduce(glu, cuda_device, small, dt, 1)[:ncomp]
    for c in range(ncomp):
        want = oracle.reduce(np.ascontiguousarray(small[:, c]), 1)
        assert abs(float(got[c]) - want) <= 1e-4 * abs(want)


@pytest.mark.parametrize("dt_name,np_dtype,ncomp", [("UVec2", np.uint32, 2), ("UVec4", np.uint32, 4),
                                                    ("IVec2", np.int32, 2), ("IVec4", np.int32, 4)])
@pytest.mark.parametrize("op", [0, 1, 2, 3])
def test_reduce_integer_vectors(glu, cuda_device, oracle, dt_name, np_dtype, ncomp, op):
    n = 123_457
    data = oracle.mt19937_u32(21, n * ncomp).reshape(n, ncomp)
    if op == 1:
        data = data | np.uint32(1)
    data = data.view(np_dtype)
    got = gpu_reduce(glu, cuda_device, data, getattr(glu, "DataType_" + dt_name), op)[:ncomp]
    want = oracle.reduce(data, op)
    assert [int(x) for x in got] == [int(x) for x in want]


def test_reduce_float_full_size_2_28(glu, cuda_device):
    """BASELINE configs[4]: Reduce(Float, Sum / Min / Max) over 2^28 uniform floats in [-1, 1).  Min / Max bit-exact
    against the device's own exact min / max (torch = plumbing; min / max are order-independent), Sum within
    1e-5 * sum|x| of a float64 accumulation of the same data — the tolerance convention of
    test/reduce_tests.cpp:72 (absolute 0.1 on sums of ~1e2) scaled to the size."""
    import torch

    n = 1 << 28
    g = torch.Generator(device=cuda_device).manual_seed(17)
    data = torch.rand(n, dtype=torch.float32, device=cuda_device, generator=g) * 2.0 - 1.0
    want_sum = float(data.sum(dtype=torch.float64).item())
    abs_sum = float(data.abs().sum(dtype=torch.float64).item())
    want_min, want_max = data.min(), data.max()
    for op, want in ((glu.ReduceOperator_Min, want_min), (glu.ReduceOperator_Max, want_max)):
        work = data.clone()
        glu.Reduce(glu.DataType_Float, op)(work, n)
        torch.cuda.synchronize()
        assert work[0].view(torch.int32).item() == want.view(torch.int32).item(), (int(op), work[0].item(), want.item())
    work = data.clone()
    glu.Reduce(glu.DataType_Float, glu.ReduceOperator_Sum)(work, n)
    torch.cuda.synchronize()
    got = float(work[0].item())
    tol = 1e-5 * abs_sum
    assert abs(got - want_sum) <= tol, (got, want_sum, tol)
    # run-to-run determinism at size (fixed combination order)
    again = data.clone()
    glu.Reduce(glu.DataType_Float, glu.ReduceOperator_Sum)(again, n)
    torch.cuda.synchronize()
    assert again[0].view(torch.int32).item() == work[0].view(torch.int32).item()


def test_reduce_float_sum_is_deterministic(glu, cuda_device):
    rng = np.random.default_rng(3)
    data = rng.uniform(-1.0, 1.0, size=1 << 22).astype(np.float32)
    a = gpu_reduce(glu, cuda_device, data, glu.DataType_Float, 0)[0]
    b = gpu_reduce(glu, cuda_device, data, glu.DataType_Float, 0)[0]
    assert a.tobytes() == b.tobytes()


def test_reduce_full_size_2_28(glu, cuda_device, oracle):
    # BASELINE config 2: Reduce(Uint, Sum) over 2^28 uint32, values [0,100) from Random(1)-like stream and
    # full-range values; checked on the device against torch (int64 accumulation, plumbing only) and, for
    # the leading 2^24 slice, against the CPU oracle.
    import torch

    n = 1 << 28
    g = torch.Generator(device=cuda_device).manual_seed(1)
    t = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=cuda_device, generator=g)
    want = int(t.sum(dtype=torch.int64).item()) & 0xFFFFFFFF
    head = to_host(t[: 1 << 24], np.uint32).copy()
    buf = t.clone()
    glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)(buf, n)
    assert int(buf[0].item()) & 0xFFFFFFFF == want
    buf = t[: 1 << 24].clone()
    glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)(buf, 1 << 24)
    assert int(buf[0].item()) & 0xFFFFFFFF == oracle.reduce(head, oracle.OP_SUM)
    mx = t.clone()
    glu.Reduce(glu.DataType_Int, glu.ReduceOperator_Max)(mx, n)
    assert int(mx[0].item()) == int(t.max().item())


def test_reduce_host_entry_point(glu, cuda_device, oracle):
    data = oracle.mt19937_u32(9, 1_000_003)
    want = oracle.reduce(data, oracle.OP_SUM)
    buf = data.copy()
    glu.reduce_host(buf, buf.size, glu.DataType_Uint, glu.ReduceOperator_Sum)
    assert int(buf[0]) == want
