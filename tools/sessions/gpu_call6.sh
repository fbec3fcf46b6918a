#!/bin/bash
# GPU-box session (1 GPU), short: ncu launch list of the bench command (final round-1 code), key-only full capture,
# skewed-input timings.
set -u
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-side-metrics > $OUT/ncu_launches.log 2>&1
for dist in zero ent16 ent16hi zipf; do
  ( timeout 60 python tools/quick_bench.py --log2n 28 --what sort --dist $dist --reps 5 2>&1 | tail -2 ) >> $OUT/skew.log
done
timeout 100 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 1 \
    -o $OUT/onesweep_keys_only python tools/quick_bench.py --log2n 28 --what ex --reps 1 > $OUT/ncu_keys_only.log 2>&1
ls -la $OUT; cat $OUT/skew.log; tail -3 $OUT/ncu_launches.log
