#!/bin/bash
# Round 2, GPU call ZH (1 GPU): the whole -m gpu suite on the final binary (small_sort_kernel included).
set -u
OUT=gpurun_out/r02zh
mkdir -p $OUT
( timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest.log
cat $OUT/pytest.log
