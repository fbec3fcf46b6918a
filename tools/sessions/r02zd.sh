#!/bin/bash
# Round 2, GPU call ZD (1 GPU): digit offset requested before the polling loop — parity subset and timing.
set -u
OUT=gpurun_out/r02zd
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_sort_gpu.py tests/test_sort_ex_gpu.py tests/test_sort_segmented_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -k "not beyond and not both_kernel and not other_kernel" 2>&1 | tail -3 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python tools/quick_bench.py --what sort --reps 15 2>&1 | grep -E "^sort|histogram" | head -2 ) > $OUT/sweep.log
cat $OUT/sweep.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -1 ) > $OUT/seg.log; cat $OUT/seg.log
