#!/bin/bash
# One GPU-box session: parity tests, quick timings, bench.py, ncu launch list and full captures.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 300 python tools/quick_bench.py --log2n 28 2>&1 | tail -8 ) > $OUT/quick28.log
( timeout 300 python tools/quick_bench.py --log2n 20 --what sort 2>&1 | tail -3 ) > $OUT/quick20.log
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > $OUT/bench.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 ) > $OUT/bench_ref.log
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-side-metrics > $OUT/ncu_launches.log 2>&1
# full captures: one onesweep pass, the scan, the reduce, the histogram (2^28)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 2 \
    -o $OUT/onesweep python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_onesweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_b32|reduce|histogram' -s 3 -c 3 \
    -o $OUT/scan_reduce python tools/quick_bench.py --log2n 28 --what scan,reduce --reps 1 > $OUT/ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 \
    -o $OUT/histogram python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_hist.log 2>&1
ls -la $OUT
cat $OUT/pytest.log $OUT/smoke.log $OUT/quick28.log $OUT/quick20.log $OUT/bench.log $OUT/bench_ref.log
