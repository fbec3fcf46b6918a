#!/bin/bash
# Multi-GPU session, bench first: usage (under gpurun --gpus N): bash tools/gpu_multi2.sh N [tag]
set -u
N=${1:-2}
TAG=${2:-r01m$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error|assert" | tail -3 ) > $OUT/bench.log
( EXCHANGES=${EXCHANGES:-p2p} timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    tools/dist_phases.py 2>&1 | grep -E "^world|Error|error" | tail -8 ) > $OUT/phases.log
cat $OUT/bench.log $OUT/phases.log
