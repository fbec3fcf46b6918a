#!/bin/bash
# Round 2, GPU call C: chain fix — timing of both kernel forms, segmented sort parity, full GPU suite.
set -u
OUT=gpurun_out/r02c
mkdir -p $OUT
for c in 8 9 10 12 13 15 17; do
  echo "== GLU_SORT_CONFIG=$c" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for c in 8 10 15; do
  echo "== GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 (no look-back: timing only)" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  for d in zero zipf; do
    echo "== GLU_SORT_CONFIG=$c dist=$d" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c timeout 60 python tools/quick_bench.py --log2n 28 --what sort --dist $d --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
for c in 10 15; do
  for r in 4 108; do
    echo "== GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
  for l in 8 32; do
    echo "== GLU_SORT_CONFIG=$c GLU_SORT_RING_TILES_PER_CTA=$l" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c GLU_SORT_RING_TILES_PER_CTA=$l timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
cat $OUT/sweep.log
( timeout 300 python -m pytest tests/test_sort_segmented_gpu.py -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_seg.log
cat $OUT/pytest_seg.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_full.log
cat $OUT/pytest_full.log
