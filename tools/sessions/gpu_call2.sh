#!/bin/bash
# GPU-box session (1 GPU): parity tests, smoke, bench.py (both arms), skewed-input timings, ncu launch list and a
# full capture of the onesweep pass.  usage (under gpurun): bash tools/gpu_call2.sh [tag]
set -u
TAG=${1:-r01h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > $OUT/pytest.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 ) > $OUT/bench_ref.log
( timeout 900 python bench.py 2>&1 | tail -3 ) > $OUT/bench.log
for dist in uniform zero ent16 ent16hi zipf; do
  ( timeout 200 python tools/quick_bench.py --log2n 28 --what sort --dist $dist --reps 5 2>&1 | tail -2 ) >> $OUT/skew.log
done
( timeout 300 python tools/quick_bench.py --log2n 28 --what scan,reduce 2>&1 | tail -4 ) > $OUT/quick28.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-side-metrics > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 1 \
    -o $OUT/onesweep python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_onesweep.log 2>&1
ls -la $OUT
cat $OUT/pytest.log $OUT/smoke.log $OUT/skew.log $OUT/quick28.log $OUT/bench_ref.log $OUT/bench.log
