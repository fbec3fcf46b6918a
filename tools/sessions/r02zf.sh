#!/bin/bash
# Round 2, GPU call ZF (1 GPU): SEG pass with both prefix rows requested together and looked at late — parity, timing.
set -u
OUT=gpurun_out/r02zf
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_sort_segmented_gpu.py -m gpu -x -q -k "not other_kernel" 2>&1 | tail -3 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -1 ) > $OUT/seg.log; cat $OUT/seg.log
