#!/bin/bash
# Round 2 multi-GPU session: usage (under gpurun --gpus N): bash tools/r02m.sh N [tag] [steps]
set -u
N=${1:-2}
TAG=${2:-r02m$N}
STEPS=${3:-10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( timeout 120 $RUN --master-port 29701 tools/nvlink_probe.py 2>&1 | grep -E "^\{|Error|error" | tail -3 ) > $OUT/nvlink.log
cat $OUT/nvlink.log
( GLU_TEST_WORLD=$N timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -k "distributed_world" 2>&1 | tail -15 ) > $OUT/pytest_world.log
cat $OUT/pytest_world.log
( EXCHANGES=p2p timeout 120 $RUN --master-port 29711 tools/dist_phases.py 2>&1 | grep -E "^world|Error|error" | tail -8 ) > $OUT/phases.log
cat $OUT/phases.log
for mode in serial pipeline; do
  ( GLU_BENCH_MODE=$mode timeout 300 $RUN --master-port 29712 bench.py --gpus $N --steps $STEPS --warmup 3 --no-side-metrics 2>&1 \
      | grep -E "^\{|Error|error|assert|Traceback" | tail -4 ) > $OUT/bench_$mode.log
  cat $OUT/bench_$mode.log
done
( timeout 600 $RUN --master-port 29713 bench.py --gpus $N --steps $STEPS --warmup 3 2>&1 \
    | grep -E "^\{|Error|error|assert|Traceback" | tail -6 ) > $OUT/bench_full.log
cat $OUT/bench_full.log
( timeout 200 $RUN --master-port 29714 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep -E "^\{|Error|error" | tail -2 ) > $OUT/bench_ref.log
cat $OUT/bench_ref.log
