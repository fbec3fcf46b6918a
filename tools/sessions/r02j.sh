#!/bin/bash
# Round 2, GPU call J (2 GPUs): world-2 parity incl. the "dma" exchange style, then the style / lane sweep at 2^28 pairs per GPU.
export SWEEPS="GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=staged
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma GLU_PIPE_LANES=2
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma GLU_PIPE_PRIORITY=none
GLU_BENCH_MODE=serial GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma"
bash tools/gpu_multi_session.sh 2 r02j 10 pytest,sweep
