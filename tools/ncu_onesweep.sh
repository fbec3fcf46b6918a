#!/bin/bash
# One ncu --set full capture of a single onesweep pass at 2^28 pairs.  usage: bash tools/ncu_onesweep.sh tag [VAR=value ...]
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 5 -c 1 \
    -o $OUT/onesweep python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_onesweep.log 2>&1
tail -3 $OUT/ncu_onesweep.log
ls -la $OUT
