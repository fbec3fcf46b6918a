#!/bin/bash
# Round 2, GPU call W (1 GPU): C++ runner (new segmented case), segmented parity after the histogram revert, its timing.
set -u
OUT=gpurun_out/r02w
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_cpp_runner_gpu.py tests/test_sort_segmented_gpu.py -m gpu -x -q 2>&1 | tail -6 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -2 ) > $OUT/seg_bench.log
cat $OUT/seg_bench.log
