"""CPU oracle for the glu hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker or as the timed CPU baseline.  The
product (``gl-radix-sort_b200`` / ``libglu_b200.so``) never imports it.

ctypes wrapper over ``oracle/libglu_oracle.so`` (built by ``make -C oracle`` from
``oracle/glu_oracle.cpp``; every function there cites the reference file:line it restates).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libglu_oracle.so")

# glu/Reduce.hpp:42-48
OP_SUM, OP_MUL, OP_MIN, OP_MAX = 0, 1, 2, 3
# glu/data_types.hpp:8-22
(DT_FLOAT, DT_DOUBLE, DT_INT, DT_UINT, DT_VEC2, DT_VEC4, DT_DVEC2, DT_DVEC4, DT_UVEC2, DT_UVEC4, DT_IVEC2,
 DT_IVEC4) = range(12)

_DT_NUMPY = {
    DT_FLOAT: (np.float32, 1), DT_DOUBLE: (np.float64, 1), DT_INT: (np.int32, 1), DT_UINT: (np.uint32, 1),
    DT_VEC2: (np.float32, 2), DT_VEC4: (np.float32, 4), DT_DVEC2: (np.float64, 2), DT_DVEC4: (np.float64, 4),
    DT_UVEC2: (np.uint32, 2), DT_UVEC4: (np.uint32, 4), DT_IVEC2: (np.int32, 2), DT_IVEC4: (np.int32, 4),
}


def build(force: bool = False) -> str:
    """Compile the oracle with the PATH g++ (the image's $CXX wrapper has no libgomp)."""
    src = os.path.join(_HERE, "glu_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        base = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-shared", "-o", _LIB_PATH, src]
        r = subprocess.run(base[:3] + ["-fopenmp"] + base[3:], capture_output=True, text=True)
        if r.returncode != 0:  # no OpenMP runtime: single-threaded oracle
            subprocess.run(base, check=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        sz = ctypes.c_size_t
        L.glu_oracle_version.restype = ctypes.c_int
        L.glu_oracle_max_threads.restype = ctypes.c_int
        L.glu_oracle_random_u32.argtypes = [ctypes.c_uint64, sz, ctypes.c_uint32, ctypes.c_uint32, u32p]
        L.glu_oracle_mt19937_u32.argtypes = [ctypes.c_uint32, sz, u32p]
        L.glu_oracle_reduce_u32.argtypes = [u32p, sz, ctypes.c_int]
        L.glu_oracle_reduce_u32.restype = ctypes.c_uint32
        L.glu_oracle_reduce_i32.argtypes = [ctypes.c_void_p, sz, ctypes.c_int]
        L.glu_oracle_reduce_i32.restype = ctypes.c_int32
        L.glu_oracle_reduce_f32.argtypes = [ctypes.c_void_p, sz, ctypes.c_int]
        L.glu_oracle_reduce_f32.restype = ctypes.c_double
        L.glu_oracle_reduce_f64.argtypes = [ctypes.c_void_p, sz, ctypes.c_int]
        L.glu_oracle_reduce_f64.restype = ctypes.c_double
        L.glu_oracle_sum_abs_f32.argtypes = [ctypes.c_void_p, sz]
        L.glu_oracle_sum_abs_f32.restype = ctypes.c_double
        L.glu_oracle_exclusive_scan_u32.argtypes = [u32p, u32p, sz, sz]
        L.glu_oracle_exclusive_scan_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, sz, sz]
        L.glu_oracle_exclusive_scan_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, sz, sz]
        L.glu_oracle_stable_sort_pairs.argtypes = [u32p, u32p, sz, sz, ctypes.c_int]
        L.glu_oracle_stable_sort_ex.argtypes = [u32p, u32p, sz, ctypes.c_uint, ctypes.c_uint, ctypes.c_int]
        L.glu_oracle_time_stable_sort_pairs.argtypes = [u32p, u32p, sz, ctypes.c_int]
        L.glu_oracle_time_stable_sort_pairs.restype = ctypes.c_double
        L.glu_oracle_lsd_sort_pairs.argtypes = [u32p, u32p, sz, sz]
        L.glu_oracle_reduce_glsl.argtypes = [ctypes.c_void_p, sz, ctypes.c_int, ctypes.c_int]
        L.glu_oracle_reduce_glsl.restype = ctypes.c_int
        L.glu_oracle_blelloch_scan_glsl_u32.argtypes = [u32p, sz, sz]
        L.glu_oracle_blelloch_scan_glsl_u32.restype = ctypes.c_int
        L.glu_oracle_radix_sort_glsl.argtypes = [u32p, u32p, sz, sz]
        L.glu_oracle_radix_sort_glsl.restype = ctypes.c_int
        _lib = L
    return _lib


def _u32p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))


def max_threads() -> int:
    return int(lib().glu_oracle_max_threads())


# ---- generators -------------------------------------------------------------------------------------------------

def random_u32(seed: int, n: int, lo: int, hi: int) -> np.ndarray:
    """glu::Random(seed).sample_int_vector<GLuint>(n, lo, hi) — test/util/Random.hpp:24-38."""
    out = np.empty(n, dtype=np.uint32)
    lib().glu_oracle_random_u32(seed, n, lo, hi, _u32p(out))
    return out


def mt19937_u32(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    lib().glu_oracle_mt19937_u32(seed, n, _u32p(out))
    return out


# ---- (A) std:: oracles ----------------------------------------------------------------------------------------

def reduce(data: np.ndarray, op: int):
    """std::accumulate / min_element / max_element; component-wise for (n, ncomp) arrays."""
    data = np.ascontiguousarray(data)
    if data.ndim == 2:
        return np.array([reduce(np.ascontiguousarray(data[:, c]), op) for c in range(data.shape[1])])
    n = data.shape[0]
    p = data.ctypes.data_as(ctypes.c_void_p)
    if data.dtype == np.uint32:
        return int(lib().glu_oracle_reduce_u32(_u32p(data), n, op))
    if data.dtype == np.int32:
        return int(lib().glu_oracle_reduce_i32(p, n, op))
    if data.dtype == np.float32:
        return float(lib().glu_oracle_reduce_f32(p, n, op))
    if data.dtype == np.float64:
        return float(lib().glu_oracle_reduce_f64(p, n, op))
    raise TypeError(data.dtype)


def sum_abs_f32(data: np.ndarray) -> float:
    data = np.ascontiguousarray(data, dtype=np.float32)
    return float(lib().glu_oracle_sum_abs_f32(data.ctypes.data_as(ctypes.c_void_p), data.size))


def exclusive_scan(data: np.ndarray, count: int | None = None, num_partitions: int = 1) -> np.ndarray:
    """std::exclusive_scan(..., 0) applied to each of num_partitions adjacent segments of `count`."""
    data = np.ascontiguousarray(data)
    if count is None:
        count = data.shape[0]
    assert data.shape[0] == count * num_partitions
    out = np.empty_like(data)
    if data.dtype == np.uint32:
        lib().glu_oracle_exclusive_scan_u32(_u32p(data), _u32p(out), count, num_partitions)
    elif data.dtype == np.int32:
        o = out.view(np.uint32)
        lib().glu_oracle_exclusive_scan_u32(_u32p(data.view(np.uint32)), _u32p(o), count, num_partitions)
    elif data.dtype == np.float32:
        lib().glu_oracle_exclusive_scan_f32(data.ctypes.data_as(ctypes.c_void_p),
                                            out.ctypes.data_as(ctypes.c_void_p), count, num_partitions)
    elif data.dtype == np.float64:
        lib().glu_oracle_exclusive_scan_f64(data.ctypes.data_as(ctypes.c_void_p),
                                            out.ctypes.data_as(ctypes.c_void_p), count, num_partitions)
    else:
        raise TypeError(data.dtype)
    return out


def stable_sort_pairs(keys: np.ndarray, vals: np.ndarray, num_steps: int = 0, threads: int = 1):
    """std::stable_sort of (key,val) pairs by key -> (sorted_keys, sorted_vals)."""
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = np.array(vals, dtype=np.uint32, copy=True)
    lib().glu_oracle_stable_sort_pairs(_u32p(k), _u32p(v), k.size, num_steps, threads)
    return k, v


def stable_sort_ex(keys: np.ndarray, vals: np.ndarray | None, begin_bit: int = 0, end_bit: int = 32,
                   descending: bool = False):
    """std::stable_sort comparing key bits [begin_bit, end_bit) only, ascending or descending; vals may be None."""
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = None if vals is None else np.array(vals, dtype=np.uint32, copy=True)
    lib().glu_oracle_stable_sort_ex(_u32p(k), None if v is None else _u32p(v), k.size, begin_bit, end_bit,
                                    1 if descending else 0)
    return k, v


def stable_sort_wide(keys: np.ndarray, vals: np.ndarray | None, descending: bool = False):
    """Oracle of glu_radix_sort_wide (beyond the reference, parity unpinned by it): stable sort of uint32 / uint64 keys,
    ascending or descending, the rows of `vals` (any element width) carried along.  numpy's stable argsort on the
    (complemented, for descending) keys == std::stable_sort with less<> / greater<>."""
    assert keys.dtype in (np.uint32, np.uint64) and keys.ndim == 1
    order = np.argsort(~keys if descending else keys, kind="stable")
    return keys[order], (None if vals is None else vals[order])


def time_stable_sort_pairs(keys: np.ndarray, vals: np.ndarray, threads: int = 1) -> float:
    """Seconds spent in std::stable_sort (threads==1) / __gnu_parallel::stable_sort on the pairs."""
    return float(lib().glu_oracle_time_stable_sort_pairs(_u32p(keys), _u32p(vals), keys.size, threads))


def lsd_sort_pairs(keys: np.ndarray, vals: np.ndarray, num_steps: int = 0):
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = np.array(vals, dtype=np.uint32, copy=True)
    lib().glu_oracle_lsd_sort_pairs(_u32p(k), _u32p(v), k.size, num_steps)
    return k, v


# ---- (B) shader-faithful restatements ---------------------------------------------------------------------------

def reduce_glsl(data: np.ndarray, data_type: int, op: int, count: int | None = None) -> np.ndarray:
    """glu::Reduce(data_type, op)(buffer, count) restated; returns the whole (clobbered) buffer."""
    dt, ncomp = _DT_NUMPY[data_type]
    buf = np.array(data, dtype=dt, copy=True).reshape(-1)
    if count is None:
        count = buf.size // ncomp
    rc = lib().glu_oracle_reduce_glsl(buf.ctypes.data_as(ctypes.c_void_p), count, data_type, op)
    if rc != 0:
        raise ValueError("reference would abort (GLU_CHECK_ARGUMENT)")
    return buf


def blelloch_scan_glsl(data: np.ndarray, count: int | None = None, num_partitions: int = 1) -> np.ndarray:
    buf = np.array(data, dtype=np.uint32, copy=True)
    if count is None:
        count = buf.size
    rc = lib().glu_oracle_blelloch_scan_glsl_u32(_u32p(buf), count, num_partitions)
    if rc != 0:
        raise ValueError("reference would abort (GLU_CHECK_ARGUMENT)")
    return buf


def radix_sort_glsl(keys: np.ndarray, vals: np.ndarray, num_steps: int = 0):
    """glu::RadixSort()(keys, vals, n, num_steps) restated -> (keys, vals, buffer_index_of_result)."""
    k = np.array(keys, dtype=np.uint32, copy=True)
    v = np.array(vals, dtype=np.uint32, copy=True)
    where = lib().glu_oracle_radix_sort_glsl(_u32p(k), _u32p(v), k.size, num_steps)
    return k, v, int(where)
