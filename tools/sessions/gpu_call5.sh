#!/bin/bash
# GPU-box session (1 GPU), short: bench.py (both arms) first, then the parity tests not run earlier today.
set -u
TAG=${1:-r01l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 200 python bench.py 2>&1 | tail -3 ) > $OUT/bench.log
( timeout 60 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 ) > $OUT/bench_ref.log
( timeout 150 python -m pytest tests/test_reduce_gpu.py tests/test_scan_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -6 ) > $OUT/pytest_rest.log
( timeout 40 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -3 ) > $OUT/smoke.log
( timeout 60 python tools/quick_bench.py --log2n 28 --what reducef --reps 5 2>&1 | tail -4 ) > $OUT/reducef.log
cat $OUT/bench.log $OUT/bench_ref.log $OUT/pytest_rest.log $OUT/smoke.log $OUT/reducef.log
