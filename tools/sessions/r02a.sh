#!/bin/bash
# Round 2, GPU call A: parity + timing of the persistent ring form of the onesweep pass (GLU_SORT_CONFIG 9..18).
set -u
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for c in 9 15 12; do
  echo "== pytest GLU_SORT_CONFIG=$c" >> $OUT/pytest.log
  ( GLU_SORT_CONFIG=$c timeout 400 python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "not beyond_2_30 and not without_tma" 2>&1 | tail -4 ) >> $OUT/pytest.log
done
for c in 8 9 10 11 12 13 14 15 16 17 18; do
  echo "== GLU_SORT_CONFIG=$c" >> $OUT/sweep.log
  ( GLU_SORT_CONFIG=$c timeout 120 python tools/quick_bench.py --log2n 28 --what sort --reps 10 2>&1 | tail -2 ) >> $OUT/sweep.log
done
for c in 8 10 15; do
  for o in 1; do
    echo "== GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=$o (no look-back: timing only)" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c GLU_SORT_OPTIONS=$o timeout 120 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
  for d in zero zipf; do
    echo "== GLU_SORT_CONFIG=$c dist=$d" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c timeout 120 python tools/quick_bench.py --log2n 28 --what sort --dist $d --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
for c in 10 15; do
  for r in 4 104 108; do
    echo "== GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r" >> $OUT/sweep.log
    ( GLU_SORT_CONFIG=$c GLU_SORT_CHAIN_ROWS=$r timeout 120 python tools/quick_bench.py --log2n 28 --what sort --reps 5 2>&1 | tail -2 ) >> $OUT/sweep.log
  done
done
echo "== memcheck config 9, 100k pairs" >> $OUT/sanitizer.log
( GLU_SORT_CONFIG=9 timeout 300 compute-sanitizer --tool memcheck python -c "
import __graft_entry__ as e, torch, numpy as np
glu=e.load_package()
n=100003
k=torch.randint(-(1<<31),(1<<31)-1,(n,),dtype=torch.int32,device='cuda'); v=torch.arange(n,dtype=torch.int32,device='cuda')
glu.RadixSort()(k,v,n); torch.cuda.synchronize(); print('ok')
" 2>&1 | tail -6 ) >> $OUT/sanitizer.log
cat $OUT/pytest.log $OUT/sweep.log $OUT/sanitizer.log
