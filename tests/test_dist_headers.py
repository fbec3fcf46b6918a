"""dist/*.hpp — the self-contained single-header copies of the three classes (the reference ships the same thing,
generate.py:7-38 -> dist/): up to date with include/, and each usable with no other project header on the include path."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

HEADERS = ["Reduce.hpp", "BlellochScan.hpp", "RadixSort.hpp"]


def test_dist_is_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "generate_dist.py"), "--check"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr + "\nrun `python tools/generate_dist.py`"


@pytest.mark.parametrize("header", HEADERS)
def test_each_dist_header_compiles_alone(tmp_path, header):
    tu = tmp_path / "tu.cpp"
    tu.write_text(f'#include "{header}"\nint main() {{ return 0; }}\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "dist"), "-fsyntax-only", str(tu)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    text = open(os.path.join(ROOT, "dist", header)).read()
    assert '#include "' not in text  # nothing project-local left to resolve
    assert "glu_radix_sort_u32kv" in text or header != "RadixSort.hpp"


def test_all_dist_headers_together_link_against_the_library(tmp_path, glu):
    tu = tmp_path / "all.cpp"
    tu.write_text("".join(f'#include "{h}"\n' for h in HEADERS) + """
int main(int argc, char**)
{
    if (argc > 100)
    { // never executed here (no GPU on the CPU box): only has to compile and link
        glu::Reduce reduce(glu::DataType_Uint, glu::ReduceOperator_Sum);
        glu::BlellochScan scan(glu::DataType_Uint);
        glu::RadixSort sort;
        reduce(nullptr, 1);
        scan(nullptr, 1);
        sort(nullptr, nullptr, 1);
        sort.sort_ex(nullptr, nullptr, 1, 0, 32, true);
    }
    return 0;
}
""")
    exe = tmp_path / "all"
    libdir = os.path.dirname(glu.LIB_PATH)
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "dist"), "-o", str(exe), str(tu),
                        "-L", libdir, "-lglu_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([str(exe)]).returncode == 0
