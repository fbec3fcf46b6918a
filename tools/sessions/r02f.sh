#!/bin/bash
# Round 2, GPU call F (2 GPUs): memory-traffic-only diagnostic of the pass, then the multi-GPU session.
set -u
OUT=gpurun_out/r02f
mkdir -p $OUT
for o in 0 5 9; do
  echo "== GLU_SORT_CONFIG=8 GLU_SORT_OPTIONS=$o (5: tile copied straight back with 4-byte stores, 9: 16-byte stores)" >> $OUT/diag.log
  ( GLU_SORT_CONFIG=8 GLU_SORT_OPTIONS=$o timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/diag.log
done
cat $OUT/diag.log
( timeout 300 python -m pytest tests/test_sort_segmented_gpu.py -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest_seg.log
cat $OUT/pytest_seg.log
bash tools/gpu_multi_session.sh 2 r02f 10 probe,pytest,phases,sweep
