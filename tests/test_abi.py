"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what
include/glu_b200.h declares, and rejects bad arguments with the documented status codes
(no compute is launched here — every rejected call returns before touching CUDA)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "glu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GLU_API\s+[\w\s\*]*?\b(glu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(glu):
    declared = declared_symbols()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", glu.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (glu_\w+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing undeclared leaks out of the library, and the Python binding covers the whole header
    assert exported == set(declared), sorted(exported - set(declared))
    assert set(glu.ABI) == set(declared)
    for name in declared:
        assert getattr(glu.lib, name) is not None


def test_library_is_self_contained_and_has_sm100a_code(glu):
    out = subprocess.run(["ldd", glu.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "not found" not in out
    elf = subprocess.run(["cuobjdump", "-lelf", glu.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf, elf


def test_enums_and_sizes_match_reference(glu):
    # glu/data_types.hpp:8-22 values 0..11; std430 strides
    sizes = [4, 8, 4, 4, 8, 16, 16, 32, 8, 16, 8, 16]
    for dt, sz in enumerate(sizes):
        assert glu.lib.glu_data_type_size(dt) == sz
    assert glu.lib.glu_data_type_size(12) == 0 and glu.lib.glu_data_type_size(-1) == 0
    assert [int(x) for x in glu.ReduceOperator] == [0, 1, 2, 3]  # glu/Reduce.hpp:42-48
    assert int(glu.DataType_Uint) == 3 and int(glu.DataType_IVec4) == 11
    assert glu.lib.glu_version() >= 100
    assert glu.lib.glu_status_string(0) == b"success"


def test_argument_errors_are_status_codes_not_exits(glu):
    L = glu.lib
    fake = ctypes.c_void_p(0x1000)  # never dereferenced: all of these fail validation first
    # glu/Reduce.hpp:113-114
    assert L.glu_reduce(None, 10, 3, 0, fake, 1 << 20, None) == 1
    assert L.glu_reduce(fake, 0, 3, 0, fake, 1 << 20, None) == 1
    assert L.glu_reduce(fake, 10, 12, 0, fake, 1 << 20, None) == 2
    assert L.glu_reduce(fake, 10, 3, 4, fake, 1 << 20, None) == 3  # glu/Reduce.hpp:93
    assert L.glu_reduce(ctypes.c_void_p(0x1002), 10, 3, 0, fake, 1 << 20, None) == 5
    assert L.glu_reduce(fake, 10, 3, 0, None, 0, None) == 4
    assert L.glu_reduce(fake, 1, 3, 0, None, 0, None) == 0  # count == 1: nothing to do (glu/Reduce.hpp:124)
    # glu/BlellochScan.hpp:132-135
    assert L.glu_scan_exclusive(None, 8, 1, 3, fake, 1 << 20, None) == 1
    assert L.glu_scan_exclusive(fake, 0, 1, 3, fake, 1 << 20, None) == 1
    assert L.glu_scan_exclusive(fake, 8, 0, 3, fake, 1 << 20, None) == 1
    assert L.glu_scan_exclusive(fake, 8, 1, 99, fake, 1 << 20, None) == 2
    assert L.glu_scan_exclusive(fake, 8, 1, 3, None, 0, None) == 4
    # glu/RadixSort.hpp:275-279
    assert L.glu_radix_sort_u32kv(None, fake, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_u32kv(fake, None, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_u32kv(fake, fake, 1, 0, None, 0, None) == 0  # count <= 1 is a silent no-op
    assert L.glu_radix_sort_u32kv(fake, fake, 0, 0, None, 0, None) == 0
    assert L.glu_radix_sort_u32kv(fake, fake, 8, 0, None, 0, None) == 4
    assert L.glu_radix_sort_u32kv(fake, fake, 1 << 31, 0, fake, 1 << 40, None) == 6
    # glu_radix_sort_u32_ex: values are optional, the bit range must be ordered and inside the key
    assert L.glu_radix_sort_u32_ex(None, fake, 8, 0, 32, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_u32_ex(fake, fake, 8, 9, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_u32_ex(fake, fake, 8, 0, 33, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_u32_ex(fake, None, 1, 0, 32, 0, None, 0, None) == 0   # count <= 1
    assert L.glu_radix_sort_u32_ex(fake, None, 8, 7, 7, 1, None, 0, None) == 0    # empty bit range: nothing to do
    assert L.glu_radix_sort_u32_ex(fake, None, 8, 0, 32, 0, None, 0, None) == 4
    assert L.glu_radix_sort_u32_ex(fake, None, 1 << 31, 0, 32, 0, fake, 1 << 40, None) == 6
    assert L.glu_radix_sort_u32_ex(ctypes.c_void_p(0x1002), None, 8, 0, 32, 0, fake, 1 << 20, None) == 5
    # glu_radix_sort_wide: element widths, value pointer <-> value_bytes, alignment to the element size
    assert L.glu_radix_sort_wide(None, 8, fake, 4, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_wide(fake, 3, fake, 4, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_wide(fake, 8, fake, 12, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_wide(fake, 8, None, 4, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_wide(fake, 8, fake, 0, 8, 0, fake, 1 << 20, None) == 1
    assert L.glu_radix_sort_wide(fake, 8, fake, 8, 1, 0, None, 0, None) == 0  # count <= 1
    assert L.glu_radix_sort_wide(ctypes.c_void_p(0x1004), 8, None, 0, 8, 0, fake, 1 << 20, None) == 5
    assert L.glu_radix_sort_wide(fake, 8, ctypes.c_void_p(0x1008), 16, 8, 0, fake, 1 << 20, None) == 5
    assert L.glu_radix_sort_wide(fake, 8, fake, 8, 8, 0, None, 0, None) == 4
    assert L.glu_radix_sort_wide(fake, 8, None, 0, 1 << 31, 0, fake, 1 << 40, None) == 6


def test_tmp_size_queries(glu):
    L = glu.lib
    assert L.glu_reduce_tmp_bytes(1 << 28, 3) >= 256
    assert L.glu_reduce_tmp_bytes(10, 99) == 0
    assert L.glu_scan_exclusive_tmp_bytes(1 << 28, 1, 3) >= (1 << 28) // 16384 * 8  # one 64-bit word per tile
    assert L.glu_scan_exclusive_tmp_bytes(0, 1, 3) == 0
    n = 1 << 20
    need = L.glu_radix_sort_u32kv_tmp_bytes(n)
    assert 2 * 4 * n <= need <= 2 * 4 * n + (8 << 20)  # two scratch arrays + O(tiles) control words
    assert L.glu_radix_sort_u32kv_tmp_bytes(1 << 31) == 0
    # key-only sorts need one scratch array, not two
    kv, ko = L.glu_radix_sort_u32_ex_tmp_bytes(n, 1), L.glu_radix_sort_u32_ex_tmp_bytes(n, 0)
    assert 2 * 4 * n <= kv <= 2 * 4 * n + (8 << 20) and 4 * n <= ko <= 4 * n + (8 << 20)
    assert L.glu_radix_sort_u32_ex_tmp_bytes(1 << 31, 0) == 0
    # wide sort: index + key word + permuted keys + permuted values + the 32-bit sort's own scratch
    w = L.glu_radix_sort_wide_tmp_bytes(n, 8, 16)
    assert w >= 4 * n + 4 * n + 8 * n + 16 * n + kv and w <= 32 * n + kv + (1 << 20)
    assert L.glu_radix_sort_wide_tmp_bytes(n, 4, 4) == kv and L.glu_radix_sort_wide_tmp_bytes(n, 4, 0) == ko
    assert L.glu_radix_sort_wide_tmp_bytes(n, 8, 3) == 0 and L.glu_radix_sort_wide_tmp_bytes(n, 2, 4) == 0
    assert L.glu_radix_sort_wide_tmp_bytes(1 << 31, 8, 0) == 0


def test_python_mirror_validates_like_the_reference(glu):
    with pytest.raises(glu.GluError):
        glu.Reduce(99, glu.ReduceOperator_Sum)
    with pytest.raises(glu.GluError):
        glu.Reduce(glu.DataType_Uint, 7)
    with pytest.raises(glu.GluError):
        glu.BlellochScan(-1)
    r = glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)
    with pytest.raises(glu.GluError):
        r(0, 10)  # "Invalid buffer"
    with pytest.raises(glu.GluError):
        r(0x1000, 0)  # "Count must be greater than zero"
    s = glu.RadixSort()
    with pytest.raises(glu.GluError):
        s(0, 0x1000, 10)
    s(0x1000, 0x1000, 1)  # no-op, never touches the device
    with pytest.raises(glu.GluError):
        s.sort_ex(0, None, 10)
    with pytest.raises(glu.GluError):
        s.sort_ex(0x1000, None, 10, 8, 4)
    s.sort_ex(0x1000, None, 10, 5, 5)  # empty bit range: no-op


def test_no_product_import_of_oracle(glu):
    """The product must not route through the oracle (or any CPU fallback): no import / include / dlopen of
    anything under oracle/, and the shared library neither links nor references it."""
    bad = re.compile(r"^\s*(import|from)\s+oracle|#\s*include\s*[<\"][^>\"]*oracle|glu_oracle_\w+\s*\(|libglu_oracle|dlopen",
                     re.M)
    roots = [os.path.join(ROOT, "gl-radix-sort_b200"), os.path.join(ROOT, "include")]
    for root in roots:
        for dirpath, _, files in os.walk(root):
            if os.path.basename(dirpath) == "build":
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(text), os.path.join(dirpath, f)
    syms = subprocess.run(["nm", "-D", glu.LIB_PATH], capture_output=True, text=True).stdout
    assert "glu_oracle" not in syms


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/glu_b200.h must compile as C99 (what a cgo / JNI / ctypes-cffi binding feeds it to),
    with no C++ and no CUDA or torch types in the signatures."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "c.c"
    src.write_text('#include "glu_b200.h"\nint main(void) { return GLU_SUCCESS; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(ROOT, "include", "glu_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)   # declarations only, comments may name what a handle is
    for banned in ("torch", "at::", "cudaStream_t", "cudaEvent_t", "std::"):
        assert banned not in code, banned
