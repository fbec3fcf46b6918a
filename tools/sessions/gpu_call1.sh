#!/bin/bash
# GPU-box session (1 GPU): parity tests, L2-prefetch sweep, skewed-input timings, bench.py (both arms), ncu launch
# list and full captures with the best prefetch distance.  usage (under gpurun): bash tools/gpu_call1.sh [tag]
set -u
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > $OUT/pytest.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $OUT/smoke.log
for d in 0 74 148 296 444 888; do
  echo "== GLU_SORT_PREFETCH=$d" >> $OUT/prefetch_sweep.log
  ( GLU_SORT_PREFETCH=$d timeout 120 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/prefetch_sweep.log
done
BEST=$(python - $OUT/prefetch_sweep.log <<'PY'
import re, sys
best, best_ms, cur = 0, 1e9, None
for line in open(sys.argv[1]):
    m = re.match(r"== GLU_SORT_PREFETCH=(\d+)", line)
    if m:
        cur = int(m.group(1))
    m = re.search(r"median ([0-9.]+) ms", line)
    if m and cur is not None and float(m.group(1)) < best_ms - 0.01:
        best, best_ms = cur, float(m.group(1))
print(best)
PY
)
echo "best prefetch distance: $BEST" | tee -a $OUT/prefetch_sweep.log
export GLU_SORT_PREFETCH=$BEST
( GLU_SORT_PREFETCH=$BEST timeout 300 python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "not beyond_2_30" 2>&1 | tail -3 ) > $OUT/pytest_prefetch.log
for dist in uniform zero ent16 ent16hi zipf; do
  ( timeout 200 python tools/quick_bench.py --log2n 28 --what sort --dist $dist --reps 5 2>&1 | tail -2 ) >> $OUT/skew.log
done
( timeout 300 python tools/quick_bench.py --log2n 28 --what scan,reduce 2>&1 | tail -4 ) > $OUT/quick28.log
( timeout 300 python tools/quick_bench.py --log2n 20 --what sort 2>&1 | tail -2 ) > $OUT/quick20.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 ) > $OUT/bench_ref.log
( timeout 900 python bench.py 2>&1 | tail -3 ) > $OUT/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-side-metrics > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep -s 4 -c 2 \
    -o $OUT/onesweep python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_onesweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_b32|reduce' -s 3 -c 2 \
    -o $OUT/scan_reduce python tools/quick_bench.py --log2n 28 --what scan,reduce --reps 1 > $OUT/ncu_scan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:histogram -s 1 -c 1 \
    -o $OUT/histogram python tools/quick_bench.py --log2n 28 --what sort --reps 1 > $OUT/ncu_hist.log 2>&1
ls -la $OUT
cat $OUT/pytest.log $OUT/smoke.log $OUT/prefetch_sweep.log $OUT/pytest_prefetch.log $OUT/skew.log $OUT/quick28.log $OUT/quick20.log $OUT/bench_ref.log $OUT/bench.log
