#!/bin/bash
# Round 2, GPU call L (N GPUs, default 4): world-N parity with the flag-signalled "dma" exchange, then pipeline timelines.
set -u
N=${1:-4}
OUT=gpurun_out/r02l
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( GLU_TEST_WORLD=$N timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -k "distributed_world" 2>&1 | tail -15 ) > $OUT/pytest_world.log
cat $OUT/pytest_world.log
for v in "GLU_PIPE_LANES=2" "GLU_PIPE_LANES=3" "GLU_PIPE_LANES=2 GLU_DIST_DMA_SYNC=nccl" "GLU_PIPE_LANES=3 GLU_PIPE_COPY_STREAM=0"; do
  echo "== $v" >> $OUT/timeline.log
  ( env $v timeout 200 $RUN --master-port 29721 tools/pipeline_timeline.py --jobs 8 < /dev/null 2>&1 | grep -E "^world|^job|^ +[0-9]|Error|error|Traceback" | tail -14 ) >> $OUT/timeline.log
done
cat $OUT/timeline.log
