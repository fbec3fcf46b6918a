#!/bin/bash
# Round 2, GPU call U (1 GPU): scan with the producer drawing tickets one stage ahead — parity, then A/B timing.
set -u
OUT=gpurun_out/r02u
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -k "scan" 2>&1 | tail -4 ) > $OUT/pytest.log
cat $OUT/pytest.log
for v in "GLU_SCAN_TICKET_AHEAD=0 GLU_SCAN_EARLY_CHAIN=0" "GLU_SCAN_TICKET_AHEAD=1 GLU_SCAN_EARLY_CHAIN=0" "GLU_SCAN_TICKET_AHEAD=1 GLU_SCAN_EARLY_CHAIN=1" "GLU_SCAN_TICKET_AHEAD=1 GLU_SCAN_EARLY_CHAIN=1 GLU_SCAN_CONFIG=10" "GLU_SCAN_TICKET_AHEAD=1 GLU_SCAN_EARLY_CHAIN=1 GLU_SCAN_CONFIG=8"; do
  echo "== $v" >> $OUT/scan.log
  ( env $v timeout 100 python tools/quick_bench.py --what scan --reps 20 2>&1 | grep "^scan" ) >> $OUT/scan.log
done
cat $OUT/scan.log
