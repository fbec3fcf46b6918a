#!/bin/bash
# Round 2, GPU call ZA (1 GPU): the late look-back as the plain pass's default — sort parity, timing, C++ runner.
set -u
OUT=gpurun_out/r02za
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_sort_ex_gpu.py tests/test_sort_wide_gpu.py tests/test_sort_segmented_gpu.py tests/test_multigpu_gpu.py tests/test_cpp_runner_gpu.py -m gpu -x -q 2>&1 | tail -4 ) > $OUT/pytest.log
cat $OUT/pytest.log
for v in "GLU_SORT_OPTIONS=0" "GLU_SORT_OPTIONS=16"; do
  echo "== $v" >> $OUT/sweep.log
  ( env $v timeout 120 python tools/quick_bench.py --what sort --reps 15 2>&1 | grep -E "^sort|histogram" | head -2 ) >> $OUT/sweep.log
done
cat $OUT/sweep.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -1 ) > $OUT/seg.log; cat $OUT/seg.log
