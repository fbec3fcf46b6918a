#!/bin/bash
# Round 2, GPU call I (1 GPU): parity of the runs variant of the segmented sort and of the "dma" exchange style at
# world 1, bench.py (e2e through the depth-3 host queue), ticket-vs-blockIdx tile ids, the README table.
set -u
OUT=gpurun_out/r02i
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > $OUT/smi.txt
( timeout 600 python -m pytest tests/test_sort_segmented_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest.log
cat $OUT/pytest.log
( timeout 400 python bench.py --steps 20 --warmup 3 2>&1 | grep -E "^\{|Error|error|assert|Traceback" | tail -3 ) > $OUT/bench.log
cat $OUT/bench.log
for o in 0 2; do
  echo "== GLU_SORT_OPTIONS=$o (2: tile ids from an atomic ticket)" >> $OUT/ticket.log
  ( GLU_SORT_OPTIONS=$o timeout 120 python tools/quick_bench.py --what sort --reps 10 2>&1 | grep -E "^sort|histogram|Error" | head -4 ) >> $OUT/ticket.log
done
cat $OUT/ticket.log
( timeout 600 python tools/readme_table.py > $OUT/readme_table.md 2> $OUT/readme_table.err; tail -3 $OUT/readme_table.err; head -45 $OUT/readme_table.md )
