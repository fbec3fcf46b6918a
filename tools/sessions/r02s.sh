#!/bin/bash
# Round 2, GPU call S (1 GPU): the scan against an in-place elementwise stream of the same size, scan tile-shape sweep.
set -u
OUT=gpurun_out/r02s
mkdir -p $OUT
( timeout 200 python tools/quick_bench.py --what scan,reduce --reps 20 2>&1 | tail -5 ) > $OUT/scan.log
for c in 6 7 8 9 10 11; do
  echo "== GLU_SCAN_CONFIG=$c" >> $OUT/scan.log
  ( GLU_SCAN_CONFIG=$c timeout 100 python tools/quick_bench.py --what scan --reps 20 2>&1 | grep "^scan" ) >> $OUT/scan.log
done
cat $OUT/scan.log
