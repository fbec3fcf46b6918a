#!/bin/bash
# Round 2, GPU call Q (1 GPU): the whole -m gpu suite and smoke() on the final code, C++ runner included.
set -u
OUT=gpurun_out/r02q
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -3 ) > $OUT/smoke.log
cat $OUT/smoke.log
( timeout 120 python tools/seg_bench.py 2>&1 | tail -2 ) > $OUT/seg_bench.log
cat $OUT/seg_bench.log
