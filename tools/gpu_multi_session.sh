#!/bin/bash
# Round 2 multi-GPU session: usage (under gpurun --gpus N): bash tools/gpu_multi_session.sh N [tag] [steps] [what]
#   what: comma list of  probe,pcie,pytest,phases,sweep,full,ref   (default: all but pcie)
#   SWEEPS (environment): the sweep's variants, one per line
set -u
N=${1:-2}
TAG=${2:-r02m$N}
STEPS=${3:-10}
WHAT=${4:-probe,pytest,phases,sweep,full,ref}
OUT=gpurun_out/$TAG
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
has() { case ",$WHAT," in *",$1,"*) return 0;; *) return 1;; esac; }
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if has probe; then
  ( timeout 120 $RUN --master-port 29701 tools/nvlink_probe.py 2>&1 | grep -E "^\{|Error|error" | tail -3 ) > $OUT/nvlink.log
  cat $OUT/nvlink.log
fi
if has pcie; then
  ( timeout 120 $RUN --master-port 29702 tools/pcie_probe.py 2>&1 | grep -E "^\{|Error|error" | tail -3 ) > $OUT/pcie.log
  cat $OUT/pcie.log
fi
if has pytest; then
  ( GLU_TEST_WORLD=$N timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -k "distributed_world" 2>&1 | tail -15 ) > $OUT/pytest_world.log
  cat $OUT/pytest_world.log
fi
if has phases; then
  for local in full segmented; do
    ( GLU_DIST_LOCAL=$local EXCHANGES=p2p timeout 120 $RUN --master-port 29711 tools/dist_phases.py 2>&1 | grep -E "^world|Error|error" | tail -4 | sed "s/^/[$local] /" ) >> $OUT/phases.log
  done
  cat $OUT/phases.log
fi
if has sweep; then
  # one variant per line in $SWEEPS (environment assignments for bench.py); default: the round's comparison set
  DEFAULT_SWEEPS="GLU_BENCH_MODE=serial GLU_DIST_LOCAL=full
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=staged
GLU_BENCH_MODE=pipeline GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma
GLU_BENCH_MODE=serial GLU_DIST_LOCAL=segmented GLU_DIST_EXCHANGE_STYLE=dma"
  echo "${SWEEPS:-$DEFAULT_SWEEPS}" | while IFS= read -r v; do
    [ -z "$v" ] && continue
    echo "== $v" >> $OUT/sweep.log
    ( env $v timeout 200 $RUN --master-port 29712 bench.py --gpus $N --steps $STEPS --warmup 3 --no-side-metrics < /dev/null 2>&1 \
        | grep -E "^\{|Error|error|assert|Traceback" | tail -3 \
        | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        r = d['roofline']
        print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'onesweep_ms': r['ms_per_launch'], 'launches': r['launches'], 'hist_ms': r['histogram_ms_per_launch'], 'exchange_ms': r['partition_exchange_ms_per_launch'], 'verified': d['verified']['ok'], 'clocks': d['clocks']['sm_mhz']}))
    except Exception:
        print(l.strip()[:600])
" ) >> $OUT/sweep.log
  done
  cat $OUT/sweep.log
fi
if has full; then
  ( timeout 900 $RUN --master-port 29713 bench.py --gpus $N --steps $STEPS --warmup 3 2>&1 \
      | grep -E "^\{|Error|error|assert|Traceback" | tail -6 ) > $OUT/bench_full.log
  cat $OUT/bench_full.log
fi
if has ref; then
  ( timeout 200 $RUN --master-port 29714 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep -E "^\{|Error|error" | tail -2 ) > $OUT/bench_ref.log
  cat $OUT/bench_ref.log
fi
