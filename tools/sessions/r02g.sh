#!/bin/bash
# Round 2, GPU call G (2 GPUs): staged exchange with local bypass; full bench.py at N = 2 (e2e, sharded reduce / scan, configs[3]).
set -u
python - <<'PY'
import re
p='tools/gpu_multi_session.sh'
s=open(p).read()
start=s.index('  for v in "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=full"')
end=s.index('    echo "== $v" >> $OUT/sweep.log')
s=s[:start]+'''  for v in "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=full" \\
           "GLU_BENCH_MODE=serial GLU_DIST_LOCAL=segmented" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=full GLU_PIPE_PRIORITY=x" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_PIPE_PRIORITY=x" \\
           "GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_PIPE_PRIORITY=s"; do
'''+s[end:]
open(p,'w').write(s)
PY
bash tools/gpu_multi_session.sh 2 r02g 10 pytest,sweep,full,ref
