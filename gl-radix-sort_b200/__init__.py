"""gl-radix-sort_b200 — B200-native (sm_100a) replacement for loryruta/gl-radix-sort's hot path.

Python mirror of the reference's three classes over the C ABI in ``include/glu_b200.h``:

    glu::Reduce(DataType, ReduceOperator)(buffer, count)                     glu/Reduce.hpp:62,111
    glu::BlellochScan(DataType)(buffer, count, num_partitions=1)             glu/BlellochScan.hpp:91,130
    glu::RadixSort()(key_buffer, val_buffer, count, num_steps=0)             glu/RadixSort.hpp:205,273
    glu::RadixSort::prepare_internal_buffers(count)                          glu/RadixSort.hpp:237

A ``buffer`` is a CUDA ``torch.Tensor`` (its ``data_ptr()`` plays the role of the GL buffer handle) or a
raw device pointer (``int``).  PyTorch is plumbing only: device memory for the scratch the reference
classes own, and the current stream.  All compute happens in ``libglu_b200.so`` (hand-written CUDA);
there is NO fallback — importing this package without the built library raises ImportError, and every
call needs a CUDA device.

The directory name contains a hyphen, so load it with ``__graft_entry__.load_package()`` (or
``importlib``) — it registers itself as ``gl_radix_sort_b200``.
"""
from __future__ import annotations

import ctypes
import enum
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglu_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C gl-radix-sort_b200`). There is no CPU / PyTorch fallback for the glu hot path.")

_lib = ctypes.CDLL(LIB_PATH)

_sz = ctypes.c_size_t
_vp = ctypes.c_void_p
_int = ctypes.c_int

# name -> (restype, argtypes); must list every symbol include/glu_b200.h declares (tests/test_abi.py checks)
ABI = {
    "glu_version": (_int, []),
    "glu_status_string": (ctypes.c_char_p, [_int]),
    "glu_last_cuda_error": (ctypes.c_char_p, []),
    "glu_data_type_size": (_sz, [_int]),
    "glu_kernel_launch_count": (ctypes.c_uint64, []),
    "glu_profile_enable": (_int, [_int]),
    "glu_profile_collect": (_int, [_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "glu_reduce_tmp_bytes": (_sz, [_sz, _int]),
    "glu_reduce": (_int, [_vp, _sz, _int, _int, _vp, _sz, _vp]),
    "glu_scan_exclusive_tmp_bytes": (_sz, [_sz, _sz, _int]),
    "glu_scan_exclusive": (_int, [_vp, _sz, _sz, _int, _vp, _sz, _vp]),
    "glu_radix_sort_u32kv_tmp_bytes": (_sz, [_sz]),
    "glu_radix_sort_u32kv": (_int, [_vp, _vp, _sz, _sz, _vp, _sz, _vp]),
    "glu_radix_sort_u32_ex_tmp_bytes": (_sz, [_sz, _int]),
    "glu_radix_sort_u32_ex": (_int, [_vp, _vp, _sz, ctypes.c_uint, ctypes.c_uint, _int, _vp, _sz, _vp]),
    "glu_radix_sort_wide_tmp_bytes": (_sz, [_sz, _sz, _sz]),
    "glu_radix_sort_wide": (_int, [_vp, _sz, _vp, _sz, _sz, _int, _vp, _sz, _vp]),
    "glu_reduce_into": (_int, [_vp, _sz, _int, _int, _vp, _vp, _sz, _vp]),
    "glu_scan_exclusive_init": (_int, [_vp, _sz, _sz, _int, _vp, _vp, _sz, _vp]),
    "glu_radix_histogram_u32": (_int, [_vp, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _vp]),
    "glu_radix_partition_u32kv_tmp_bytes": (_sz, [_sz]),
    "glu_radix_partition_u32kv": (_int, [_vp, _vp, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _sz, _vp]),
    "glu_radix_partition_by_dest_u32kv": (_int, [_vp, _vp, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _vp, _sz,
                                                 _vp]),
    "glu_radix_sort_u32kv_dyn": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _sz, _vp]),
    "glu_radix_partition_by_dest_u32kv_dyn": (_int, [_vp, _vp, _vp, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _vp,
                                                     _sz, _vp]),
    "glu_radix_exchange_plan": (_int, [_vp, _int, _int, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "glu_radix_partition_u32kv_dyn": (_int, [_vp, _vp, _vp, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _sz, _vp]),
    "glu_radix_exchange_plan_buckets": (_int, [_vp, _int, _int, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "glu_radix_exchange_stage_tables": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "glu_radix_exchange_copy_u32kv": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "glu_radix_sort_segment_tile": (_sz, []),
    "glu_radix_sort_u32kv_segmented_tmp_bytes": (_sz, [_sz]),
    "glu_radix_sort_u32kv_segmented": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, ctypes.c_uint, ctypes.c_uint, _vp, _sz,
                                              _vp, ctypes.POINTER(_int)]),
    "glu_radix_sort_u32kv_segmented_runs": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, ctypes.c_uint, ctypes.c_uint, _vp,
                                                   _sz, _vp, _sz, _vp, ctypes.POINTER(_int)]),
    "glu_ipc_get_handle": (_int, [_vp, ctypes.c_char_p]),
    "glu_ipc_open_handle": (_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "glu_ipc_close_handle": (_int, [_vp]),
    "glu_reduce_host": (_int, [_vp, _sz, _int, _int]),
    "glu_scan_exclusive_host": (_int, [_vp, _sz, _sz, _int]),
    "glu_radix_sort_u32kv_host": (_int, [_vp, _vp, _sz, _sz]),
    "glu_host_sort_queue_create": (_int, [ctypes.POINTER(_vp), _sz, _int]),
    "glu_host_sort_queue_submit": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "glu_host_sort_queue_wait": (_int, [_vp]),
    "glu_host_sort_queue_destroy": (_int, [_vp]),
    "glu_device_count": (_int, [ctypes.POINTER(_int)]),
    "glu_set_device": (_int, [_int]),
    "glu_device_info": (_int, [_int, ctypes.c_char_p, _sz, ctypes.POINTER(_int), ctypes.POINTER(_int),
                               ctypes.POINTER(_int), ctypes.POINTER(_sz), ctypes.POINTER(_int)]),
    "glu_malloc": (_int, [ctypes.POINTER(_vp), _sz]),
    "glu_free": (_int, [_vp]),
    "glu_malloc_host": (_int, [ctypes.POINTER(_vp), _sz]),
    "glu_free_host": (_int, [_vp]),
    "glu_memcpy_h2d": (_int, [_vp, _vp, _sz, _vp]),
    "glu_memcpy_d2h": (_int, [_vp, _vp, _sz, _vp]),
    "glu_memcpy_d2d": (_int, [_vp, _vp, _sz, _vp]),
    "glu_memset_u32": (_int, [_vp, ctypes.c_uint32, _sz, _vp]),
    "glu_signal_peers_u32": (_int, [_vp, _int, ctypes.c_uint32, _vp]),
    "glu_stream_wait_flags_u32": (_int, [_vp, _int, _int, ctypes.c_uint32, _vp]),
    "glu_stream_create": (_int, [ctypes.POINTER(_vp)]),
    "glu_stream_destroy": (_int, [_vp]),
    "glu_stream_synchronize": (_int, [_vp]),
    "glu_event_create": (_int, [ctypes.POINTER(_vp)]),
    "glu_event_destroy": (_int, [_vp]),
    "glu_event_record": (_int, [_vp, _vp]),
    "glu_event_synchronize": (_int, [_vp]),
    "glu_event_elapsed_ms": (_int, [ctypes.POINTER(ctypes.c_float), _vp, _vp]),
}
for _name, (_res, _args) in ABI.items():
    _fn = getattr(_lib, _name)  # AttributeError here == the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args

lib = _lib


class DataType(enum.IntEnum):
    """glu/data_types.hpp:8-22"""
    Float = 0
    Double = 1
    Int = 2
    Uint = 3
    Vec2 = 4
    Vec4 = 5
    DVec2 = 6
    DVec4 = 7
    UVec2 = 8
    UVec4 = 9
    IVec2 = 10
    IVec4 = 11


class ReduceOperator(enum.IntEnum):
    """glu/Reduce.hpp:42-48"""
    Sum = 0
    Mul = 1
    Min = 2
    Max = 3


# reference-style aliases
DataType_Float, DataType_Double, DataType_Int, DataType_Uint = DataType.Float, DataType.Double, DataType.Int, DataType.Uint
DataType_Vec2, DataType_Vec4, DataType_DVec2, DataType_DVec4 = DataType.Vec2, DataType.Vec4, DataType.DVec2, DataType.DVec4
DataType_UVec2, DataType_UVec4, DataType_IVec2, DataType_IVec4 = DataType.UVec2, DataType.UVec4, DataType.IVec2, DataType.IVec4
ReduceOperator_Sum, ReduceOperator_Mul, ReduceOperator_Min, ReduceOperator_Max = (
    ReduceOperator.Sum, ReduceOperator.Mul, ReduceOperator.Min, ReduceOperator.Max)


class GluError(RuntimeError):
    """Raised where the reference's GLU_CHECK_ARGUMENT / GLU_FAIL would print and exit(1) (glu/errors.hpp:8-18)."""

    def __init__(self, status: int, where: str):
        self.status = status
        msg = _lib.glu_status_string(status).decode()
        if status == 7:
            msg += ": " + _lib.glu_last_cuda_error().decode()
        super().__init__(f"{where}: {msg}")


def check(status: int, where: str) -> None:
    if status != 0:
        raise GluError(status, where)


def kernel_launch_count() -> int:
    return int(_lib.glu_kernel_launch_count())


KERNEL_REDUCE, KERNEL_SCAN, KERNEL_SORT_HISTOGRAM, KERNEL_SORT_ONESWEEP, KERNEL_SORT_PARTITION = 0, 1, 2, 3, 4


def profile_enable(on: bool) -> None:
    """Bracket every hot-path kernel launch with CUDA events on its stream (bench.py's roofline)."""
    check(_lib.glu_profile_enable(1 if on else 0), "profile_enable")


def profile_collect(kernel_id: int):
    """(total_ms, launches) of one kernel family since the last collect."""
    ms, n = ctypes.c_double(0), ctypes.c_uint64(0)
    check(_lib.glu_profile_collect(kernel_id, ctypes.byref(ms), ctypes.byref(n)), "profile_collect")
    return float(ms.value), int(n.value)


def data_type_size(data_type: int) -> int:
    return int(_lib.glu_data_type_size(int(data_type)))


def _ptr_and_device(buffer):
    """Device pointer (+ torch device or None) of a buffer argument."""
    if isinstance(buffer, int):
        return buffer, None
    if hasattr(buffer, "data_ptr"):
        if not buffer.is_cuda:
            raise GluError(1, "buffer must live on a CUDA device (no CPU path exists)")
        if not buffer.is_contiguous():
            raise GluError(1, "buffer must be contiguous")
        return buffer.data_ptr(), buffer.device
    raise TypeError(f"unsupported buffer type {type(buffer)!r}")


def _current_stream(device) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)


class _Scratch:
    """Grow-only device scratch owned by an operator object (glu/RadixSort.hpp:193-200, :237-271)."""

    def __init__(self):
        self._buf = None

    def ensure(self, nbytes: int, device):
        import torch

        if self._buf is None or self._buf.numel() < nbytes or self._buf.device != device:
            self._buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            if os.environ.get("GLU_VERBOSE"):
                print(f"[glu] scratch reallocated to: {nbytes}")
        return self._buf.data_ptr(), self._buf.numel()


def _device_of(dev):
    import torch

    return dev if dev is not None else torch.device("cuda", torch.cuda.current_device())


class Reduce:
    """glu::Reduce — in-place reduction, result in element 0 (glu/Reduce.hpp:51-135)."""

    def __init__(self, data_type, operator_):
        if int(data_type) not in set(int(d) for d in DataType):
            raise GluError(2, f"Invalid data type: {int(data_type)}")
        if int(operator_) not in (0, 1, 2, 3):
            raise GluError(3, f"Invalid reduction operator: {int(operator_)}")  # glu/Reduce.hpp:93
        self.data_type = DataType(int(data_type))
        self.operator = ReduceOperator(int(operator_))
        self._scratch = _Scratch()

    def __call__(self, buffer, count: int, stream: int | None = None) -> None:
        ptr, dev = _ptr_and_device(buffer)
        if not ptr:
            raise GluError(1, "Invalid buffer")
        if count <= 0:
            raise GluError(1, "Count must be greater than zero")
        dev = _device_of(dev)
        need = int(_lib.glu_reduce_tmp_bytes(count, int(self.data_type)))
        tmp, tmp_bytes = self._scratch.ensure(need, dev)
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_reduce(ptr, count, int(self.data_type), int(self.operator), tmp, tmp_bytes, st), "Reduce")


class BlellochScan:
    """glu::BlellochScan — in-place exclusive prefix sum over num_partitions adjacent segments
    (glu/BlellochScan.hpp:79-139).  Any count is accepted (the reference requires a power of two)."""

    def __init__(self, data_type):
        if int(data_type) not in set(int(d) for d in DataType):
            raise GluError(2, f"Invalid data type: {int(data_type)}")
        self.data_type = DataType(int(data_type))
        self._scratch = _Scratch()

    def __call__(self, buffer, count: int, num_partitions: int = 1, stream: int | None = None) -> None:
        ptr, dev = _ptr_and_device(buffer)
        if not ptr:
            raise GluError(1, "Invalid buffer")
        if count <= 0:
            raise GluError(1, "Count must be greater than zero")
        if num_partitions < 1:
            raise GluError(1, "Num of partitions must be >= 1")
        dev = _device_of(dev)
        need = int(_lib.glu_scan_exclusive_tmp_bytes(count, num_partitions, int(self.data_type)))
        tmp, tmp_bytes = self._scratch.ensure(need, dev)
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_scan_exclusive(ptr, count, num_partitions, int(self.data_type), tmp, tmp_bytes, st),
              "BlellochScan")


class RadixSort:
    """glu::RadixSort — stable in-place sort of uint32 (key, value) pairs by key (glu/RadixSort.hpp:186-354)."""

    def __init__(self):
        self._scratch = _Scratch()
        self._device = None

    def prepare_internal_buffers(self, count: int, device=None) -> None:
        """glu/RadixSort.hpp:237-271 — grow-only pre-sizing of the internal scratch."""
        dev = _device_of(device if device is not None else self._device)
        need = int(_lib.glu_radix_sort_u32kv_tmp_bytes(count))
        if need == 0:
            raise GluError(6, "RadixSort")
        self._scratch.ensure(need, dev)
        self._device = dev

    def __call__(self, key_buffer, val_buffer, count: int, num_steps: int = 0, stream: int | None = None) -> None:
        kptr, kdev = _ptr_and_device(key_buffer)
        vptr, vdev = _ptr_and_device(val_buffer)
        if not kptr:
            raise GluError(1, "Invalid key buffer")
        if not vptr:
            raise GluError(1, "Invalid value buffer")
        if count <= 1:
            return  # glu/RadixSort.hpp:278-279
        dev = _device_of(kdev if kdev is not None else vdev)
        self.prepare_internal_buffers(count, dev)
        tmp, tmp_bytes = self._scratch.ensure(0, dev)
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_radix_sort_u32kv(kptr, vptr, count, num_steps, tmp, tmp_bytes, st), "RadixSort")

    def sort_ex(self, key_buffer, val_buffer, count: int, begin_bit: int = 0, end_bit: int = 32,
                descending: bool = False, stream: int | None = None) -> None:
        """glu_radix_sort_u32_ex: `val_buffer=None` sorts keys only; only key bits [begin_bit, end_bit) take part;
        `descending` puts the largest key first.  Always stable (equal keys keep their input order)."""
        kptr, kdev = _ptr_and_device(key_buffer)
        vptr, vdev = _ptr_and_device(val_buffer) if val_buffer is not None else (None, None)
        if not kptr:
            raise GluError(1, "Invalid key buffer")
        if val_buffer is not None and not vptr:
            raise GluError(1, "Invalid value buffer")
        if not (0 <= begin_bit <= end_bit <= 32):
            raise GluError(1, "RadixSort.sort_ex: need 0 <= begin_bit <= end_bit <= 32")
        if count <= 1 or begin_bit == end_bit:
            return
        dev = _device_of(kdev if kdev is not None else vdev)
        need = int(_lib.glu_radix_sort_u32_ex_tmp_bytes(count, 1 if vptr else 0))
        if need == 0:
            raise GluError(6, "RadixSort.sort_ex")
        tmp, tmp_bytes = self._scratch.ensure(need, dev)
        self._device = dev
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_radix_sort_u32_ex(kptr, vptr, count, begin_bit, end_bit, 1 if descending else 0, tmp, tmp_bytes,
                                         st), "RadixSort.sort_ex")

    def sort_wide(self, key_buffer, val_buffer, count: int, key_bytes: int = 8, value_bytes: int = 4,
                  descending: bool = False, stream: int | None = None) -> None:
        """glu_radix_sort_wide: 4- or 8-byte unsigned keys with no (`val_buffer=None`, `value_bytes=0`), 4-, 8- or
        16-byte values.  Stable, in place."""
        kptr, kdev = _ptr_and_device(key_buffer)
        vptr, vdev = _ptr_and_device(val_buffer) if val_buffer is not None else (None, None)
        if not kptr:
            raise GluError(1, "Invalid key buffer")
        if (val_buffer is None) != (value_bytes == 0) or (val_buffer is not None and not vptr):
            raise GluError(1, "Invalid value buffer / value_bytes")
        if key_bytes not in (4, 8) or value_bytes not in (0, 4, 8, 16):
            raise GluError(1, "RadixSort.sort_wide: key_bytes must be 4 or 8, value_bytes 0, 4, 8 or 16")
        if count <= 1:
            return
        dev = _device_of(kdev if kdev is not None else vdev)
        need = int(_lib.glu_radix_sort_wide_tmp_bytes(count, key_bytes, value_bytes))
        if need == 0:
            raise GluError(6, "RadixSort.sort_wide")
        tmp, tmp_bytes = self._scratch.ensure(need, dev)
        self._device = dev
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_radix_sort_wide(kptr, key_bytes, vptr, value_bytes, count, 1 if descending else 0, tmp, tmp_bytes,
                                       st), "RadixSort.sort_wide")

    def sort_segmented(self, keys_a, vals_a, keys_b, vals_b, seg_count_buffer, num_segments: int, max_tiles: int,
                       begin_bit: int = 0, end_bit: int = 32, stream: int | None = None, runs_buffer=None,
                       num_runs: int = 0) -> bool:
        """glu_radix_sort_u32kv_segmented: `num_segments` independent stable sorts by key bits [begin_bit, end_bit) in one
        set of launches.  Input in arrays A, segment s at element first_tile[s] * segment_tile() (first_tile = exclusive
        scan of ceil(count / tile)); output compact.  Returns True when the result is in arrays B, False for A.
        With `runs_buffer` (5 x (num_runs + 1) uint32, see glu_radix_sort_u32kv_segmented_runs) the input is a sequence of
        tile-aligned runs placed anywhere in arrays A."""
        ptrs = []
        dev = None
        for buf in (keys_a, vals_a, keys_b, vals_b, seg_count_buffer):
            ptr, d = _ptr_and_device(buf)
            if not ptr:
                raise GluError(1, "Invalid buffer")
            dev = dev if dev is not None else d
            ptrs.append(ptr)
        dev = _device_of(dev)
        need = int(_lib.glu_radix_sort_u32kv_segmented_tmp_bytes(max_tiles))
        if need == 0:
            raise GluError(6, "RadixSort.sort_segmented")
        tmp, tmp_bytes = self._scratch.ensure(need, dev)
        self._device = dev
        st = _current_stream(dev) if stream is None else stream
        in_b = ctypes.c_int(0)
        if runs_buffer is not None:
            rptr, _ = _ptr_and_device(runs_buffer)
            if not rptr:
                raise GluError(1, "Invalid run table")
            check(_lib.glu_radix_sort_u32kv_segmented_runs(ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], num_segments,
                                                           max_tiles, begin_bit, end_bit, rptr, num_runs, tmp, tmp_bytes,
                                                           st, ctypes.byref(in_b)), "RadixSort.sort_segmented (runs)")
            return bool(in_b.value)
        check(_lib.glu_radix_sort_u32kv_segmented(ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], num_segments, max_tiles,
                                                  begin_bit, end_bit, tmp, tmp_bytes, st, ctypes.byref(in_b)),
              "RadixSort.sort_segmented")
        return bool(in_b.value)

    def sort_device_count(self, key_buffer, val_buffer, count_buffer, max_count: int, num_steps: int = 0,
                          stream: int | None = None) -> None:
        """The same sort with the number of pairs read from device memory (`count_buffer`: one uint32, <= max_count)
        when the kernels run — glu_radix_sort_u32kv_dyn, a building block of the multi-GPU sort."""
        kptr, kdev = _ptr_and_device(key_buffer)
        vptr, vdev = _ptr_and_device(val_buffer)
        cptr, _ = _ptr_and_device(count_buffer)
        if not kptr or not vptr or not cptr:
            raise GluError(1, "Invalid key / value / count buffer")
        if max_count <= 1:
            return
        dev = _device_of(kdev if kdev is not None else vdev)
        self.prepare_internal_buffers(max_count, dev)
        tmp, tmp_bytes = self._scratch.ensure(0, dev)
        st = _current_stream(dev) if stream is None else stream
        check(_lib.glu_radix_sort_u32kv_dyn(kptr, vptr, cptr, max_count, num_steps, tmp, tmp_bytes, st),
              "RadixSort.sort_device_count")


# ---- host-buffer entry points (numpy arrays; upload + hot path + download inside the call) ------------------------

def _np_ptr(a):
    import numpy as np

    if not isinstance(a, np.ndarray) or not a.flags["C_CONTIGUOUS"] or not a.flags["WRITEABLE"]:
        raise GluError(1, "host buffers must be writable C-contiguous numpy arrays")
    return a.ctypes.data


def reduce_host(data, count: int, data_type, operator_) -> None:
    check(_lib.glu_reduce_host(_np_ptr(data), count, int(data_type), int(operator_)), "reduce_host")


def scan_exclusive_host(data, count: int, num_partitions: int, data_type) -> None:
    check(_lib.glu_scan_exclusive_host(_np_ptr(data), count, num_partitions, int(data_type)), "scan_exclusive_host")


def radix_sort_u32kv_host(keys, vals, count: int | None = None, num_steps: int = 0) -> None:
    if count is None:
        count = keys.size
    check(_lib.glu_radix_sort_u32kv_host(_np_ptr(keys), _np_ptr(vals), count, num_steps), "radix_sort_u32kv_host")


class HostSortQueue:
    """glu_host_sort_queue_*: up to `depth` host-buffer sorts in flight, so that the upload of one job overlaps the sort
    and the download of the previous one.  The numpy arrays passed to submit() must stay alive (and should be pinned)
    until wait() returns; the results land in them in place."""

    def __init__(self, max_count: int, depth: int = 2):
        self._q = _vp()
        check(_lib.glu_host_sort_queue_create(ctypes.byref(self._q), max_count, depth), "host_sort_queue_create")
        self._keepalive = []

    def submit(self, keys, vals, count: int | None = None, num_steps: int = 0) -> None:
        if count is None:
            count = keys.size
        self._keepalive.append((keys, vals))
        check(_lib.glu_host_sort_queue_submit(self._q, _np_ptr(keys), _np_ptr(vals), count, num_steps),
              "host_sort_queue_submit")

    def wait(self) -> None:
        check(_lib.glu_host_sort_queue_wait(self._q), "host_sort_queue_wait")
        self._keepalive.clear()

    def close(self) -> None:
        if self._q:
            check(_lib.glu_host_sort_queue_destroy(self._q), "host_sort_queue_destroy")
            self._q = _vp()
            self._keepalive.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


from . import distributed  # noqa: E402  (multi-GPU composition; imports torch lazily)
from .distributed import (DistributedBlellochScan, DistributedRadixSort, DistributedReduce,  # noqa: E402,F401
                          DistributedSortPipeline)
