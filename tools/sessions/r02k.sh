#!/bin/bash
# Round 2, GPU call K (8 GPUs): world-8 parity, the exchange-style / lane sweep at 2^28 pairs per GPU, full bench.py
# (e2e, sharded reduce / scan, configs[3] at 2^30 pairs per GPU = 2^33 pairs) with the new defaults (dma, two lanes).
export SWEEPS="GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma GLU_PIPE_LANES=3
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=staged"
bash tools/gpu_multi_session.sh 8 r02k 10 pytest,sweep,full
