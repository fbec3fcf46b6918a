#!/bin/bash
# Round 2, GPU call Z (1 GPU): look-back asked for after the value scatter (GLU_SORT_OPTIONS=16) — parity, then A/B.
set -u
OUT=gpurun_out/r02z
mkdir -p $OUT
( GLU_SORT_OPTIONS=16 timeout 600 python -m pytest tests/test_sort_gpu.py tests/test_sort_segmented_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -k "not full_size and not beyond and not both_kernel and not other_kernel and not distributed_world" 2>&1 | tail -4 ) > $OUT/pytest.log
cat $OUT/pytest.log
for v in "GLU_SORT_OPTIONS=0" "GLU_SORT_OPTIONS=16" "GLU_SORT_OPTIONS=0 GLU_SORT_CHAIN_ROWS=4" "GLU_SORT_OPTIONS=16 GLU_SORT_CHAIN_ROWS=4" "GLU_SORT_OPTIONS=16" "GLU_SORT_OPTIONS=0"; do
  echo "== $v" >> $OUT/sweep.log
  ( env $v timeout 120 python tools/quick_bench.py --what sort --reps 15 2>&1 | grep -E "^sort|histogram" | head -2 ) >> $OUT/sweep.log
done
cat $OUT/sweep.log
( GLU_SORT_OPTIONS=16 timeout 120 python tools/seg_bench.py 2>&1 | tail -1 ) > $OUT/seg.log; cat $OUT/seg.log
