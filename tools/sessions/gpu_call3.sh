#!/bin/bash
# GPU-box session (1 GPU): the new glu_radix_sort_u32_ex tests first, then the sort / C++ runner tests, sort_ex timings,
# a short bench.  usage (under gpurun): bash tools/gpu_call3.sh [tag]
set -u
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 600 python -m pytest tests/test_sort_ex_gpu.py -m gpu -q --durations=5 2>&1 | tail -40 ) > $OUT/pytest_ex.log
( timeout 600 python -m pytest tests/test_cpp_runner_gpu.py tests/test_sort_gpu.py -m gpu -x -q --durations=5 2>&1 | tail -20 ) > $OUT/pytest_sort.log
( timeout 300 python tools/quick_bench.py --log2n 28 --what ex --reps 5 2>&1 | tail -8 ) > $OUT/quick_ex.log
( timeout 200 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -3 ) > $OUT/quick28.log
cat $OUT/pytest_ex.log $OUT/pytest_sort.log $OUT/quick_ex.log $OUT/quick28.log
