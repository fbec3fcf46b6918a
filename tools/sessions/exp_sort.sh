#!/bin/bash
# Tuning sweep for the sort on the GPU box: each argument is a space-separated list of VAR=value settings.
# usage: bash tools/exp_sort.sh tag "GLU_SORT_CONFIG=1" "GLU_SORT_CONFIG=1 GLU_SORT_CHAIN_ROWS=16" ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout ${PYTEST_TIMEOUT:-240} python -m pytest tests/test_sort_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -k "${PYTEST_K:-not beyond_2_30}" --durations=5 2>&1 | tail -12 ) > $OUT/pytest_sort.log
cat $OUT/pytest_sort.log
for cfg in "$@"; do
  echo "== $cfg" | tee -a $OUT/sweep.log
  ( env $cfg timeout 120 python tools/quick_bench.py --log2n ${LOG2N:-28} --what sort --reps ${REPS:-10} 2>&1 | tail -3 ) | tee -a $OUT/sweep.log
done
