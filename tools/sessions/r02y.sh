#!/bin/bash
# Round 2, GPU call Y (1 GPU): final validation — the whole -m gpu suite, smoke(), the default bench line and the
# reference arm exactly as the driver runs them.
set -u
OUT=gpurun_out/r02y
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 ) > $OUT/smoke.log
cat $OUT/smoke.log
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | grep -E "^\{|Error|Traceback" | tail -2 ) > $OUT/bench_ref.log
cut -c1-400 $OUT/bench_ref.log
( timeout 600 python bench.py 2>&1 | grep -E "^\{|Error|error|assert|Traceback" | tail -3 ) > $OUT/bench.log
cat $OUT/bench.log
