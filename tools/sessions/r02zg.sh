#!/bin/bash
# Round 2, GPU call ZG (1 GPU): small_sort_kernel (<= 2048 pairs: one CTA, one launch) — parity, then the README table.
set -u
OUT=gpurun_out/r02zg
mkdir -p $OUT
( timeout 150 python -m pytest tests/test_sort_gpu.py tests/test_sort_ex_gpu.py tests/test_cpp_runner_gpu.py -m gpu -x -q \
    -k "small_inputs or reference_cases or heavy_duplicates or num_steps or unaligned or object_reuse or count_0 or host_entry or env_extra7 or env_extra8 or sort_ex or cpp" 2>&1 | tail -4 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 60 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 ) > $OUT/smoke.log
cat $OUT/smoke.log
( timeout 100 python tools/readme_table.py > $OUT/readme_table.md 2> $OUT/readme_table.err; grep "Radix sort" $OUT/readme_table.md | head -8 )
