#!/bin/bash
# Round 2, GPU call E: ring publishes before its look-back; r1 chain for the one-tile-per-CTA kernel; ncu capture of the ring pass.
set -u
OUT=gpurun_out/r02e
mkdir -p $OUT
for c in 8 9 10 12 13 17; do
  echo "== GLU_SORT_CONFIG=$c" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for c in 8 10 12; do
  echo "== GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 (no look-back: timing only)" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=1 timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for v in "GLU_SORT_CONFIG=12 GLU_SORT_CHAIN_ROWS=4" "GLU_SORT_CONFIG=12 GLU_SORT_CHAIN_ROWS=108" "GLU_SORT_CONFIG=12 GLU_SORT_RING_CTAS_PER_SM=1" "GLU_SORT_CONFIG=12 GLU_SORT_TMA=0"; do
  echo "== $v" >> $OUT/sweep.log
  ( env $v timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
done
cat $OUT/sweep.log
( timeout 600 python -m pytest tests/test_sort_gpu.py tests/test_sort_segmented_gpu.py -m gpu -x -q -k "not beyond_2_30 and not full_size" 2>&1 | tail -8 ) > $OUT/pytest.log
cat $OUT/pytest.log
GLU_SORT_CONFIG=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_ring -s 5 -c 1 \
    -o $OUT/ring12 python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_ring12.log 2>&1
tail -3 $OUT/ncu_ring12.log
ls -la $OUT
