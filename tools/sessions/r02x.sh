#!/bin/bash
# Round 2, GPU call X (2 GPUs): the full bench line at N = 2 with the final defaults (what the driver's scaling run does).
bash tools/gpu_multi_session.sh 2 r02x 10 full,ref
