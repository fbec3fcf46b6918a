#!/bin/bash
# Round 2, GPU call ZE (1 GPU): the default bench line of the final code (late look-back in the plain pass).
set -u
OUT=gpurun_out/r02ze
mkdir -p $OUT
( timeout 600 python bench.py 2>&1 | grep -E "^\{|Error|error|assert|Traceback" | tail -3 ) > $OUT/bench.log
cut -c1-1800 $OUT/bench.log
