#!/bin/bash
# LOAD x MFIRST variants of the onesweep pass: parity (all sort tests forced onto the large-tile kernel), then timing.
OUT=gpurun_out/r01i
mkdir -p $OUT
for v in "GLU_SORT_LOAD=2 GLU_SORT_MFIRST=1" "GLU_SORT_LOAD=1 GLU_SORT_MFIRST=1" "GLU_SORT_LOAD=0 GLU_SORT_MFIRST=1" "GLU_SORT_LOAD=2 GLU_SORT_MFIRST=0" "GLU_SORT_LOAD=1 GLU_SORT_MFIRST=0"; do
  echo "== pytest $v" >> $OUT/pytest.log
  ( env $v GLU_SORT_CONFIG=8 timeout 300 python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "not beyond_2_30 and not without_tma" 2>&1 | tail -4 ) >> $OUT/pytest.log
done
for v in "GLU_SORT_LOAD=0 GLU_SORT_MFIRST=0" "GLU_SORT_LOAD=1 GLU_SORT_MFIRST=0" "GLU_SORT_LOAD=2 GLU_SORT_MFIRST=0" \
         "GLU_SORT_LOAD=0 GLU_SORT_MFIRST=1" "GLU_SORT_LOAD=1 GLU_SORT_MFIRST=1" "GLU_SORT_LOAD=2 GLU_SORT_MFIRST=1" \
         "GLU_SORT_LOAD=2 GLU_SORT_MFIRST=1 GLU_SORT_PREFETCH=0" "GLU_SORT_LOAD=1 GLU_SORT_MFIRST=1 GLU_SORT_PREFETCH=296"; do
  echo "== $v" >> $OUT/sweep.log
  ( env $v timeout 120 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for dist in zero ent16 zipf; do
  echo "== LOAD=2 MFIRST=1 $dist" >> $OUT/sweep.log
  ( GLU_SORT_LOAD=2 GLU_SORT_MFIRST=1 timeout 120 python tools/quick_bench.py --log2n 28 --what sort --dist $dist --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
done
cat $OUT/pytest.log $OUT/sweep.log
