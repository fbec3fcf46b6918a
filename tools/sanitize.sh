#!/bin/bash
# compute-sanitizer passes over small cases (run under gpurun): memcheck on the reference-style C++ suite,
# racecheck (shared-memory hazards) and synccheck on one small sort / scan / reduce.
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/small_cases.py <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
import oracle
glu = entry.load_package()
dev = torch.device("cuda", 0)
for n in (5000, 40_003):
    keys = oracle.mt19937_u32(1, n); vals = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int32)).to(dev); dv = torch.from_numpy(vals.view(np.int32)).to(dev)
    glu.RadixSort()(dk, dv, n); torch.cuda.synchronize()
    ek, ev = oracle.stable_sort_pairs(keys, vals)
    assert np.array_equal(dk.cpu().numpy().view(np.uint32), ek) and np.array_equal(dv.cpu().numpy().view(np.uint32), ev)
data = oracle.random_u32(123, 100_001, 0, 100)
dd = torch.from_numpy(data.view(np.int32)).to(dev)
glu.BlellochScan(glu.DataType_Uint)(dd, data.size); torch.cuda.synchronize()
assert np.array_equal(dd.cpu().numpy().view(np.uint32), oracle.exclusive_scan(data))
big = oracle.random_u32(5, (1 << 22) + 3, 0, 100)   # large enough for the persistent TMA scan kernel
dd = torch.from_numpy(big.view(np.int32)).to(dev)
glu.BlellochScan(glu.DataType_Uint)(dd, big.size); torch.cuda.synchronize()
assert np.array_equal(dd.cpu().numpy().view(np.uint32), oracle.exclusive_scan(big))
dd = torch.from_numpy(data.view(np.int32)).to(dev)
glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Sum)(dd, data.size)
assert int(dd[0].item()) & 0xFFFFFFFF == oracle.reduce(data, oracle.OP_SUM)
print("small cases ok")
PY
# round 2: the segmented sort, its runs form (tile map) and the peer-signalling kernels
cat > /tmp/seg_cases.py <<'PY'
import ctypes, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as entry
import oracle
import test_sort_segmented_gpu as t
glu = entry.load_package()
dev = torch.device("cuda", 0)
tile = int(glu.lib.glu_radix_sort_segment_tile())
t.run_case(glu, dev, oracle, [tile + 1, 1, 0, 3 * tile - 1, 12345, 2, tile - 1], 0, 24, seed=3)
t.run_runs_case(glu, dev, oracle, [[tile + 1, 1, 0, 2 * tile - 1], [12345, 2], [tile - 1, tile + 7, 9]], 0, 24, seed=3)
t.run_runs_case(glu, dev, oracle, [[0, 7, 0], [0, 0], [1, 0, 0, 3], [0]], 0, 8, seed=4)
flags = torch.zeros(32, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
addrs = (ctypes.c_uint64 * 4)(flags.data_ptr(), flags.data_ptr() + 4, 0, flags.data_ptr() + 12)
glu.check(glu.lib.glu_signal_peers_u32(addrs, 4, 7, st), "signal")
glu.check(glu.lib.glu_stream_wait_flags_u32(flags.data_ptr(), 4, 2, 7, st), "wait")
torch.cuda.synchronize()
assert flags[:4].tolist() == [7, 7, 0, 7]
print("segmented cases ok")
PY
( timeout 600 $CS --tool memcheck --error-exitcode 9 ./cpp_tests/glu_test RadixSort-multiple-sizes RadixSort-2048 RadixSort-segmented-and-runs BlellochScan-multiple-partitions Reduce-subgroup-fitting-size Reduce-all 2>&1 | tail -8 ) > $OUT/memcheck.log
( timeout 900 $CS --tool memcheck --error-exitcode 9 python /tmp/seg_cases.py 2>&1 | grep -E "Invalid|ERROR|SUMMARY|segmented cases|Error|at .*\.cu" | cut -c1-260 | tail -12 ) > $OUT/memcheck_seg.log
( timeout 900 $CS --tool racecheck --error-exitcode 9 python /tmp/seg_cases.py 2>&1 | grep -E "Race reported|and (Read|Write) access|hazards|ERROR|SUMMARY|segmented cases|Error" | sed -E "s/\[clone[^]]*\]//" | cut -c1-260 | sort | uniq -c | sort -rn | head -12 ) > $OUT/racecheck_seg.log
for tool in racecheck synccheck; do
  ( timeout 900 $CS --tool $tool --error-exitcode 9 python /tmp/small_cases.py 2>&1 | grep -E "Race reported|and (Read|Write) access|hazards|ERROR|SUMMARY|small cases|Error" | sed -E "s/\[clone[^]]*\]//" | cut -c1-260 ) > $OUT/$tool.log
done
tail -n 8 $OUT/memcheck.log $OUT/memcheck_seg.log $OUT/racecheck_seg.log $OUT/racecheck.log $OUT/synccheck.log
