#!/bin/bash
# Round 2, GPU call V (1 GPU): ncu --set full of the multi-GPU step's kernels on one GPU — the SEG flavour of the
# onesweep pass and the segmented histogram (tools/seg_bench.py: 32 buckets x 2^23 pairs, bits [0, 24)).
set -u
OUT=gpurun_out/r02v
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'onesweep|seg_histogram' -s 8 -c 4 \
    -o $OUT/seg python tools/seg_bench.py > $OUT/ncu_seg.log 2>&1
tail -3 $OUT/ncu_seg.log
ls -la $OUT
