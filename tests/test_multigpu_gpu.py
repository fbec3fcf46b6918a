"""GPU parity tests of the multi-GPU building blocks (glu_reduce_into, glu_scan_exclusive_init,
glu_radix_histogram_u32, glu_radix_partition_u32kv) and of gl-radix-sort_b200/distributed.py, which is run in
one process per GPU under torch.distributed.run over NCCL with as many ranks as the box has GPUs (at most 2;
with one GPU the same code runs as a world of 1, which still goes through histogram -> plan -> peer-pointer
partition -> local sort).  The oracle is the checker only."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import to_device, to_host

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stream(dev):
    import torch

    return int(torch.cuda.current_stream(dev).cuda_stream)


@pytest.mark.parametrize("n", [1, 5, 4097, 1_000_003])
@pytest.mark.parametrize("shift,bits", [(0, 8), (8, 8), (24, 8), (27, 5), (3, 1)])
def test_radix_histogram(glu, cuda_device, oracle, n, shift, bits):
    import torch

    keys = oracle.mt19937_u32(3, n + 1)
    dk = to_device(keys, cuda_device)
    hist = torch.full((256,), -1, dtype=torch.int32, device=cuda_device)
    for off in (0, 1):  # 16-byte aligned and misaligned key pointers
        glu.check(glu.lib.glu_radix_histogram_u32(dk.data_ptr() + 4 * off, n, shift, bits, hist.data_ptr(),
                                                  _stream(cuda_device)), "hist")
        want = np.bincount((keys[off:off + n] >> shift) & ((1 << bits) - 1), minlength=256)
        np.testing.assert_array_equal(to_host(hist, np.uint32), want.astype(np.uint32))


@pytest.mark.parametrize("n", [1, 33, 7679, 7680, 7681, 200_003, 3_000_017])
@pytest.mark.parametrize("shift,bits,kind", [(24, 8, "uniform"), (8, 8, "uniform"), (0, 8, "dups"), (28, 4, "uniform")])
def test_radix_partition_matches_stable_partition(glu, cuda_device, oracle, n, shift, bits, kind):
    import torch

    keys = oracle.mt19937_u32(5, n)
    if kind == "dups":
        keys = keys % np.uint32(3)
    vals = np.arange(n, dtype=np.uint32)
    digit = (keys >> shift) & ((1 << bits) - 1)
    counts = np.bincount(digit, minlength=256).astype(np.int64)
    offs = np.cumsum(counts) - counts
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    ok = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ov = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    tables = torch.from_numpy(np.concatenate([ok.data_ptr() + 4 * offs, ov.data_ptr() + 4 * offs])).to(cuda_device)
    tmp = torch.empty(int(glu.lib.glu_radix_partition_u32kv_tmp_bytes(n)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_radix_partition_u32kv(dk.data_ptr(), dv.data_ptr(), n, shift, bits, tables.data_ptr(),
                                                tables.data_ptr() + 8 * 256, tmp.data_ptr(), tmp.numel(),
                                                _stream(cuda_device)), "partition")
    order = np.argsort(digit, kind="stable")
    np.testing.assert_array_equal(to_host(ok, np.uint32), keys[order])
    np.testing.assert_array_equal(to_host(ov, np.uint32), vals[order])
    np.testing.assert_array_equal(to_host(dk, np.uint32), keys)  # inputs untouched


@pytest.mark.parametrize("n", [1, 33, 7679, 7680, 7681, 200_003, 3_000_017])
@pytest.mark.parametrize("ndest,kind", [(1, "uniform"), (2, "uniform"), (3, "dups"), (8, "uniform"), (16, "uniform")])
def test_radix_partition_by_dest_matches_stable_partition(glu, cuda_device, oracle, n, ndest, kind):
    import torch

    keys = oracle.mt19937_u32(6, n)
    if kind == "dups":
        keys = (keys % np.uint32(5)) << np.uint32(24)
    vals = np.arange(n, dtype=np.uint32)
    lut = (np.arange(256) * ndest // 256).astype(np.uint8)  # contiguous bucket ranges, like assign_buckets
    to = lut[(keys >> 24) & 0xFF]
    counts = np.bincount(to, minlength=ndest).astype(np.int64)
    offs = np.cumsum(counts) - counts
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    ok = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ov = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    tab = np.zeros(2 * 256 + 32, dtype=np.int64)
    tab[:ndest] = ok.data_ptr() + 4 * offs
    tab[256:256 + ndest] = ov.data_ptr() + 4 * offs
    tab[512:].view(np.uint8)[:] = lut
    tables = torch.from_numpy(tab).to(cuda_device)
    tmp = torch.empty(int(glu.lib.glu_radix_partition_u32kv_tmp_bytes(n)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_radix_partition_by_dest_u32kv(dk.data_ptr(), dv.data_ptr(), n, 24, 8, tables.data_ptr() + 4096,
                                                        tables.data_ptr(), tables.data_ptr() + 2048, tmp.data_ptr(),
                                                        tmp.numel(), _stream(cuda_device)), "partition_by_dest")
    order = np.argsort(to, kind="stable")
    np.testing.assert_array_equal(to_host(ok, np.uint32), keys[order])
    np.testing.assert_array_equal(to_host(ov, np.uint32), vals[order])


@pytest.mark.parametrize("n", [33, 7681, 200_003])
def test_radix_partition_by_dest_non_monotone_table_and_16_pointers(glu, cuda_device, oracle, n):
    """The destination table need not be monotone (the padding slots of the partial last tile must not depend on it),
    and only the first 16 entries of the pointer tables are read (include/glu_b200.h)."""
    import torch

    keys = oracle.mt19937_u32(9, n)
    vals = np.arange(n, dtype=np.uint32)
    ndest = 5
    lut = ((np.arange(256) * 7 + 3) % ndest).astype(np.uint8)  # scrambled: lut[255] is not the largest destination
    assert lut[255] != ndest - 1
    to = lut[(keys >> 24) & 0xFF]
    counts = np.bincount(to, minlength=ndest).astype(np.int64)
    offs = np.cumsum(counts) - counts
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    ok = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ov = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ktab = torch.from_numpy(np.concatenate([ok.data_ptr() + 4 * offs, np.zeros(16 - ndest, np.int64)])).to(cuda_device)
    vtab = torch.from_numpy(np.concatenate([ov.data_ptr() + 4 * offs, np.zeros(16 - ndest, np.int64)])).to(cuda_device)
    dlut = torch.from_numpy(lut.copy()).to(cuda_device)
    tmp = torch.empty(int(glu.lib.glu_radix_partition_u32kv_tmp_bytes(n)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_radix_partition_by_dest_u32kv(dk.data_ptr(), dv.data_ptr(), n, 24, 8, dlut.data_ptr(),
                                                        ktab.data_ptr(), vtab.data_ptr(), tmp.data_ptr(), tmp.numel(),
                                                        _stream(cuda_device)), "partition_by_dest")
    torch.cuda.synchronize()
    order = np.argsort(to, kind="stable")
    np.testing.assert_array_equal(to_host(ok, np.uint32), keys[order])
    np.testing.assert_array_equal(to_host(ov, np.uint32), vals[order])


def test_radix_partition_short_pointer_table(glu, cuda_device, oracle):
    """glu_radix_partition_u32kv with bits < 8 reads only 1 << bits pointers (include/glu_b200.h)."""
    import torch

    n, shift, bits = 100_003, 29, 3
    keys = oracle.mt19937_u32(10, n)
    vals = np.arange(n, dtype=np.uint32)
    digit = (keys >> shift) & ((1 << bits) - 1)
    counts = np.bincount(digit, minlength=1 << bits).astype(np.int64)
    offs = np.cumsum(counts) - counts
    dk, dv = to_device(keys, cuda_device), to_device(vals, cuda_device)
    ok = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ov = torch.zeros(n, dtype=torch.int32, device=cuda_device)
    ktab = torch.from_numpy(ok.data_ptr() + 4 * offs).to(cuda_device)  # exactly 8 pointers
    vtab = torch.from_numpy(ov.data_ptr() + 4 * offs).to(cuda_device)
    tmp = torch.empty(int(glu.lib.glu_radix_partition_u32kv_tmp_bytes(n)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_radix_partition_u32kv(dk.data_ptr(), dv.data_ptr(), n, shift, bits, ktab.data_ptr(),
                                                vtab.data_ptr(), tmp.data_ptr(), tmp.numel(), _stream(cuda_device)),
              "partition")
    torch.cuda.synchronize()
    order = np.argsort(digit, kind="stable")
    np.testing.assert_array_equal(to_host(ok, np.uint32), keys[order])
    np.testing.assert_array_equal(to_host(ov, np.uint32), vals[order])


def test_reduce_into_leaves_the_data_alone(glu, cuda_device, oracle):
    import torch

    data = oracle.mt19937_u32(11, 1_000_003)
    dd = to_device(data, cuda_device)
    out = torch.zeros(4, dtype=torch.int32, device=cuda_device)
    tmp = torch.empty(int(glu.lib.glu_reduce_tmp_bytes(data.size, 3)), dtype=torch.uint8, device=cuda_device)
    for op in (oracle.OP_SUM, oracle.OP_MUL, oracle.OP_MIN, oracle.OP_MAX):
        glu.check(glu.lib.glu_reduce_into(dd.data_ptr(), data.size, 3, op, out.data_ptr() + 4, tmp.data_ptr(), tmp.numel(),
                                          _stream(cuda_device)), "reduce_into")
        assert int(to_host(out, np.uint32)[1]) == oracle.reduce(data, op)
    np.testing.assert_array_equal(to_host(dd, np.uint32), data)
    glu.check(glu.lib.glu_reduce_into(dd.data_ptr() + 40, 1, 3, 0, out.data_ptr(), tmp.data_ptr(), tmp.numel(),
                                      _stream(cuda_device)), "reduce_into")
    assert int(to_host(out, np.uint32)[0]) == int(data[10])


@pytest.mark.parametrize("n", [1, 100, 8192, 70_001, (1 << 23) + 5])
def test_scan_init_uint(glu, cuda_device, oracle, n):
    import torch

    data = oracle.mt19937_u32(12, n)
    init = np.array([0xFFFFFF00], dtype=np.uint32)
    dd, di = to_device(data, cuda_device), to_device(init, cuda_device)
    tmp = torch.empty(int(glu.lib.glu_scan_exclusive_tmp_bytes(n, 1, 3)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_scan_exclusive_init(dd.data_ptr(), n, 1, 3, di.data_ptr(), tmp.data_ptr(), tmp.numel(),
                                              _stream(cuda_device)), "scan_init")
    want = (oracle.exclusive_scan(data).astype(np.uint64) + int(init[0])).astype(np.uint32)
    np.testing.assert_array_equal(to_host(dd, np.uint32), want)


def test_scan_init_partitions_and_wide_types(glu, cuda_device, oracle):
    import torch

    # every partition starts from init
    data = oracle.random_u32(123, 1000 * 37, 0, 100)
    dd, di = to_device(data, cuda_device), to_device(np.array([7], dtype=np.uint32), cuda_device)
    tmp = torch.empty(int(glu.lib.glu_scan_exclusive_tmp_bytes(1000, 37, 3)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_scan_exclusive_init(dd.data_ptr(), 1000, 37, 3, di.data_ptr(), tmp.data_ptr(), tmp.numel(),
                                              _stream(cuda_device)), "scan_init")
    np.testing.assert_array_equal(to_host(dd, np.uint32), oracle.exclusive_scan(data, 1000, 37) + np.uint32(7))
    # double (wide kernel): integers stored as doubles stay exact
    n = 50_001
    d64 = oracle.random_u32(5, n, 0, 1000).astype(np.float64)
    dd = torch.from_numpy(d64.copy()).to(cuda_device)
    di = torch.tensor([1e6], dtype=torch.float64, device=cuda_device)
    tmp = torch.empty(int(glu.lib.glu_scan_exclusive_tmp_bytes(n, 1, 1)), dtype=torch.uint8, device=cuda_device)
    glu.check(glu.lib.glu_scan_exclusive_init(dd.data_ptr(), n, 1, 1, di.data_ptr(), tmp.data_ptr(), tmp.numel(),
                                              _stream(cuda_device)), "scan_init")
    want = 1e6 + np.concatenate([[0.0], np.cumsum(d64)[:-1]])
    np.testing.assert_array_equal(dd.cpu().numpy(), want)


def test_peer_signal_and_flag_wait(glu, cuda_device):
    """glu_signal_peers_u32 / glu_stream_wait_flags_u32 on one device: the signal is ordered after the stream's earlier
    work, the wait lets the stream continue once every word (but `skip`) has reached the epoch, wrap-around included."""
    import ctypes

    import torch

    flags = torch.zeros(32, dtype=torch.int32, device=cuda_device)
    out = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    st = _stream(cuda_device)
    base = flags.data_ptr()
    for epoch in (1, 2, 0x7FFFFFFF, 0x80000001, 0xFFFFFFFF, 3):   # 3 after 0xFFFFFFFF: the counter wrapped
        addrs = (ctypes.c_uint64 * 5)(base, base + 4, 0, base + 12, base + 16)   # word 2 is never signalled
        glu.check(glu.lib.glu_signal_peers_u32(addrs, 5, epoch, st), "signal")
        glu.check(glu.lib.glu_stream_wait_flags_u32(base, 5, 2, epoch, st), "wait")
        out.add_(1)
    torch.cuda.synchronize()
    assert int(out.item()) == 6
    got = flags[:5].cpu().numpy().view(np.uint32).tolist()
    assert got == [3, 3, 0, 3, 3]
    # argument checks
    addrs = (ctypes.c_uint64 * 1)(base + 2)
    assert glu.lib.glu_signal_peers_u32(addrs, 1, 1, st) != 0          # misaligned word
    assert glu.lib.glu_signal_peers_u32(addrs, 17, 1, st) != 0         # more than 16 peers
    assert glu.lib.glu_stream_wait_flags_u32(base, 33, 0, 1, st) != 0  # more than 32 words
    assert glu.lib.glu_stream_wait_flags_u32(0, 4, 0, 1, st) != 0


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["GLU_ROOT"])
import __graft_entry__ as entry
import oracle
glu = entry.load_package()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
launches0 = glu.kernel_launch_count()

def up(a):
    return torch.from_numpy(a.view(np.int32).copy()).to(dev)

def gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out

# ---- sort: uniform 32-bit, the reference's 31-bit keys (adaptive split), heavy duplicates, 16-bit entropy
# exchange "p2p" runs twice: destination-major exchange + full local sort, and bucket-major exchange + segmented local
# sort on the key bits below the split digit (north_star's MSD split + 24-bit local sort)
for exchange in os.environ["GLU_EXCHANGES"].split(","):
  for local, style in ((("full", "staged"), ("segmented", "staged"), ("segmented", "direct"), ("segmented", "dma"))
                       if exchange == "p2p" else (("full", "staged"),)):
    os.environ["GLU_DIST_EXCHANGE_STYLE"] = style  # how the bucket-major exchange crosses NVLink (segmented only)
    sorter_fixed = glu.DistributedRadixSort(400_000, exchange=exchange, capacity_factor=2.5, local=local)
    sorter_auto = glu.DistributedRadixSort(400_000, exchange=exchange, capacity_factor=2.5, split_shift="auto", local=local)
    assert sorter_fixed.exchange == exchange, (sorter_fixed.exchange, exchange)
    assert sorter_fixed.local == local, (sorter_fixed.local, local)
    for kind, sorter in (("uniform", sorter_fixed), ("uniform", sorter_auto), ("ref31", sorter_auto), ("dups", sorter_auto),
                         ("ent16", sorter_auto), ("uniform_again", sorter_fixed)):
        n = 300_007 + 1013 * rank
        if kind.startswith("uniform"):
            keys = oracle.mt19937_u32(20 + rank, n)
        elif kind == "ref31":
            keys = oracle.random_u32(1 + rank, n, 0, 0xFFFFFFFF)
        elif kind == "dups":
            keys = oracle.random_u32(1 + rank, n, 0, 10)
        else:
            keys = oracle.mt19937_u32(20 + rank, n) & np.uint32(0xFFFF)
        sizes = gather(n)
        base = sum(sizes[:rank])
        vals = np.arange(base, base + n, dtype=np.uint32)
        dk, dv = up(keys), up(vals)
        sk, sv, m = sorter(dk, dv, n)
        torch.cuda.synchronize()
        res = gather((keys, vals, sk.cpu().numpy().view(np.uint32), sv.cpu().numpy().view(np.uint32)))
        assert np.array_equal(dk.cpu().numpy().view(np.uint32), keys), "inputs were modified"
        if rank == 0:
            allk = np.concatenate([r[0] for r in res]); allv = np.concatenate([r[1] for r in res])
            gk = np.concatenate([r[2] for r in res]); gv = np.concatenate([r[3] for r in res])
            ek, ev = oracle.stable_sort_pairs(allk, allv)
            assert np.array_equal(gk, ek), f"{exchange}/{local}/{style}/{kind}: keys differ"
            assert np.array_equal(gv, ev), f"{exchange}/{local}/{style}/{kind}: values differ (stability across ranks)"
        dist.barrier()
    sorter_fixed.close()
    sorter_auto.close()
os.environ.pop("GLU_DIST_EXCHANGE_STYLE", None)

# ---- the two-lane pipeline (DistributedSortPipeline): consecutive independent jobs, the exchange of job k+1 runs under
# the local sort of job k; every job's result must equal std::stable_sort of the concatenated inputs
for pipe_local, pipe_style in (("segmented", "dma"), ("segmented", "staged"), ("full", "staged")):
  os.environ["GLU_DIST_LOCAL"] = pipe_local
  os.environ["GLU_DIST_EXCHANGE_STYLE"] = pipe_style
  pipe = glu.DistributedSortPipeline(400_000, capacity_factor=2.5)
  assert pipe.lanes[0].local == pipe_local
  assert pipe_local == "full" or pipe.lanes[0].exchange_style == pipe_style
  jobs, pending = [], []

  def check_job(j, ticket):
      sk, sv, m = pipe.result(ticket)
      keys, vals = jobs[j][0], jobs[j][1]
      res = gather((keys, vals, sk.cpu().numpy().view(np.uint32), sv.cpu().numpy().view(np.uint32)))
      if rank == 0:
          allk = np.concatenate([r[0] for r in res]); allv = np.concatenate([r[1] for r in res])
          gk = np.concatenate([r[2] for r in res]); gv = np.concatenate([r[3] for r in res])
          ek, ev = oracle.stable_sort_pairs(allk, allv)
          assert np.array_equal(gk, ek), f"pipeline job {j}: keys differ"
          assert np.array_equal(gv, ev), f"pipeline job {j}: values differ"

  for j in range(7):
      n = 250_003 + 977 * rank + 1000 * j
      keys = oracle.mt19937_u32(60 + 10 * j + rank, n)
      if j == 3:
          keys = keys & np.uint32(0xFF0000FF)  # heavy duplicates inside every bucket
      sizes = gather(n)
      base = sum(sizes[:rank])
      vals = np.arange(base, base + n, dtype=np.uint32)
      dk, dv = up(keys), up(vals)
      jobs.append((keys, vals, dk, dv))
      pending.append((j, pipe.submit(dk, dv, n)))
      if len(pending) == pipe.num_lanes:
          check_job(*pending.pop(0))
  while pending:
      check_job(*pending.pop(0))
  pipe.close()
del os.environ["GLU_DIST_LOCAL"]
del os.environ["GLU_DIST_EXCHANGE_STYLE"]

# ---- reduce / scan sharded by contiguous ranges
n = 1_000_003 + 31 * rank
data = oracle.mt19937_u32(40 + rank, n)
alld = np.concatenate(gather(data))
for op in (oracle.OP_SUM, oracle.OP_MUL, oracle.OP_MIN, oracle.OP_MAX):
    dd = up(data)
    glu.DistributedReduce(glu.DataType_Uint, op)(dd, n)
    got = int(dd[0].item()) & 0xFFFFFFFF
    assert got == oracle.reduce(alld, op), (op, got)
f = (oracle.mt19937_u32(50 + rank, n).astype(np.float64) / 2**31 - 1.0).astype(np.float32)
allf = np.concatenate(gather(f))
for op in (oracle.OP_SUM, oracle.OP_MIN, oracle.OP_MAX):
    df = torch.from_numpy(f.copy()).to(dev)
    glu.DistributedReduce(glu.DataType_Float, op)(df, n)
    got = float(df[0].item())
    if op == oracle.OP_SUM:
        want = float(allf.astype(np.float64).sum())
        assert abs(got - want) <= 1e-5 * float(np.abs(allf.astype(np.float64)).sum()), (got, want)
    else:
        assert got == float(allf.min() if op == oracle.OP_MIN else allf.max())
dd = up(data)
glu.DistributedBlellochScan(glu.DataType_Uint)(dd, n)
sizes = gather(n)
base = sum(sizes[:rank])
want = oracle.exclusive_scan(alld)[base:base + n]
assert np.array_equal(dd.cpu().numpy().view(np.uint32), want), "sharded scan differs"
torch.cuda.synchronize()
if rank == 0:
    print("OK world", world, "launches", glu.kernel_launch_count() - launches0, "lib", glu.LIB_PATH)
dist.destroy_process_group()
'''


def _run_world(tmp_path, nproc, exchanges):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GLU_ROOT=ROOT, GLU_EXCHANGES=exchanges)
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert f"OK world {nproc}" in r.stdout


def test_distributed_world(tmp_path, cuda_device):
    import torch

    # two ranks by default; GLU_TEST_WORLD raises it on boxes with more GPUs (tools/gpu_multi_session.sh)
    nproc = min(int(os.environ.get("GLU_TEST_WORLD", "2")), torch.cuda.device_count())
    _run_world(tmp_path, nproc, "p2p,nccl")
