#!/bin/bash
# Round 2, GPU call N (1 GPU): parity after the segmented-histogram change, its timing, ncu launch list of the bench
# command and one --set full capture of the onesweep pass (final round-2 code).
set -u
OUT=gpurun_out/r02n
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_sort_segmented_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -5 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -2 ) > $OUT/seg_bench.log
cat $OUT/seg_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-side-metrics > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 \
    -o $OUT/onesweep python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_onesweep.log 2>&1
tail -2 $OUT/ncu_onesweep.log
ls -la $OUT
