#!/bin/bash
# Round 2, GPU call P (8 GPUs): the final multi-GPU code — world-8 parity, one pipeline timeline, full bench.py (e2e,
# sharded reduce / scan, configs[3] at 2^30 pairs per GPU).
set -u
OUT=gpurun_out/r02p
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 200 $RUN --master-port 29721 tools/pipeline_timeline.py --jobs 8 < /dev/null 2>&1 | grep -E "^world|^job|^ +[0-9]|Error|error|Traceback" | tail -14 ) > $OUT/timeline.log
cat $OUT/timeline.log
bash tools/gpu_multi_session.sh 8 r02p 10 pytest,full
