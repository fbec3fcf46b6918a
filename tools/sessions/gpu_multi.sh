#!/bin/bash
# Multi-GPU session: usage (under gpurun --gpus N): bash tools/gpu_multi.sh N [tag]
set -u
N=${1:-2}
TAG=${2:-r01m$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( GLU_TEST_WORLD=$N timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -k distributed_world 2>&1 | tail -15 ) > $OUT/pytest.log
( EXCHANGES=${EXCHANGES:-p2p,nccl} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 \
    tools/dist_phases.py 2>&1 | grep -E "^world|Error|error" | tail -8 ) > $OUT/phases.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tail -3 ) > $OUT/bench.log
cat $OUT/pytest.log $OUT/phases.log $OUT/bench.log
