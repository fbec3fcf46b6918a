#!/bin/bash
# Round 2, GPU call D: chain with overlapped waits — timing of both kernel forms + sort parity.
set -u
OUT=gpurun_out/r02d
mkdir -p $OUT
for c in 8 9 10 11 12 13 15 17; do
  echo "== GLU_SORT_CONFIG=$c" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for c in 8 10; do
  echo "== GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 (no look-back: timing only)" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  for d in zero zipf; do
    echo "== GLU_SORT_CONFIG=$c dist=$d" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c timeout 60 python tools/quick_bench.py --log2n 28 --what sort --dist $d --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
for v in "GLU_SORT_CONFIG=10 GLU_SORT_CHAIN_ROWS=4" "GLU_SORT_CONFIG=10 GLU_SORT_CHAIN_ROWS=108" "GLU_SORT_CONFIG=10 GLU_SORT_RING_TILES_PER_CTA=16" "GLU_SORT_CONFIG=10 GLU_SORT_RING_CTAS_PER_SM=1"; do
  echo "== $v" >> $OUT/sweep.log
  ( env $v timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
done
cat $OUT/sweep.log
( timeout 600 python -m pytest tests/test_sort_gpu.py tests/test_sort_segmented_gpu.py -m gpu -x -q -k "not beyond_2_30" 2>&1 | tail -8 ) > $OUT/pytest.log
cat $OUT/pytest.log
