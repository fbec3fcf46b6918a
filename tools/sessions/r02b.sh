#!/bin/bash
# Round 2, GPU call B: ring kernel (after the chain fix) parity + timing, the rewritten bench.py at N = 1, full GPU suite.
set -u
OUT=gpurun_out/r02b
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== baseline GLU_SORT_CONFIG=8" >> $OUT/sweep.log
( GLU_SORT_CONFIG=8 timeout 90 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
for c in 9 15; do
  echo "== pytest GLU_SORT_CONFIG=$c" >> $OUT/pytest.log
  ( GLU_SORT_CONFIG=$c timeout 240 python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "reference_cases or ragged or skewed_inputs or num_steps or heavy or unaligned or reuse" 2>&1 | tail -4 ) >> $OUT/pytest.log
done
for c in 9 10 11 12 13 14 15 16 17 18; do
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
  for r in 4 104 108; do
    echo "== GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r timeout 60 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
cat $OUT/sweep.log $OUT/pytest.log
( timeout 400 python bench.py --steps 6 --warmup 3 2>&1 | tail -3 ) > $OUT/bench.log
( timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 ) > $OUT/bench_ref.log
cat $OUT/bench.log $OUT/bench_ref.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_full.log
cat $OUT/pytest_full.log
