#!/bin/bash
# Round 2, GPU call H (8 GPUs): probes, world-8 parity, mode sweep at 2^28 pairs per GPU, full bench.py (e2e, sharded
# reduce / scan, configs[3] at 2^30 pairs per GPU = 2^33 pairs).
set -u
python - <<'PY'
p='tools/gpu_multi_session.sh'
s=open(p).read()
start=s.index('  for v in "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=full"')
end=s.index('    echo "== $v" >> $OUT/sweep.log')
s=s[:start]+'''  for v in "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=full" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=full GLU_PIPE_PRIORITY=x" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_PIPE_PRIORITY=x" \\
           "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=segmented"; do
'''+s[end:]
open(p,'w').write(s)
PY
bash tools/gpu_multi_session.sh 8 r02h 10 probe,pcie,pytest,sweep,full
