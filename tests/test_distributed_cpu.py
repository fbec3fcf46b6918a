"""CPU tests (no GPU) of the host-side logic of the multi-GPU path: bucket assignment, exchange layout, and a
world_size-2 gloo run that performs the planned all-to-all on CPU tensors.  The data movement itself is emulated
with numpy here (the product does it in glu_radix_partition_u32kv on the GPU); what is under test is that the PLAN
— destinations, offsets, split sizes, stability across ranks — yields the oracle's stable sort."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dmod(glu):
    return glu.distributed


def test_choose_split_shift(dmod):
    assert dmod.choose_split_shift(0, 0xFFFFFFFF) == 24
    assert dmod.choose_split_shift(48271, 2147483426) == 23      # the reference generator's 31-bit keys
    assert dmod.choose_split_shift(0, 0xFFFF) == 8                # 16-bit-entropy keys
    assert dmod.choose_split_shift(5, 5) == 0
    assert dmod.choose_split_shift(0x12340000, 0x1234FFFF) == 8   # constant high bits do not matter


def test_assign_buckets_balanced_and_monotone(dmod):
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        counts = rng.integers(0, 1000, size=256)
        dest = dmod.assign_buckets(counts, world)
        assert dest.min() >= 0 and dest.max() <= world - 1
        assert np.all(np.diff(dest) >= 0)
        loads = np.bincount(dest, weights=counts, minlength=world)
        assert loads.max() <= counts.sum() / world + counts.max()
    # everything in one bucket: it cannot be split, one rank takes all
    one = np.zeros(256, dtype=np.int64)
    one[17] = 12345
    assert len(set(dmod.assign_buckets(one, 8).tolist())) >= 1
    assert dmod.assign_buckets(np.zeros(256), 4).tolist() == [0] * 256


def emulate(dmod, shards_k, shards_v, shift, layout="bucket"):
    """numpy emulation of DistributedRadixSort's data movement driven by plan_exchange."""
    world = len(shards_k)
    hist_all = np.stack([np.bincount((k >> shift) & 0xFF, minlength=256) for k in shards_k])
    plan = dmod.plan_exchange(hist_all)
    recv_k = [np.zeros(int(t), dtype=np.uint32) for t in plan.recv_totals]
    recv_v = [np.zeros(int(t), dtype=np.uint32) for t in plan.recv_totals]
    written = [np.zeros(int(t), dtype=np.int32) for t in plan.recv_totals]
    for s in range(world):
        digit = (shards_k[s] >> shift) & 0xFF
        if layout == "bucket":  # 256-way pointer-table partition
            for b in range(256):
                sel = np.nonzero(digit == b)[0]  # stable partition: source order inside a bucket
                g, off = int(plan.dest[b]), int(plan.dst_offset[s][b])
                recv_k[g][off:off + sel.size] = shards_k[s][sel]
                recv_v[g][off:off + sel.size] = shards_v[s][sel]
                written[g][off:off + sel.size] += 1
        else:  # partition by destination: source order inside a destination
            to = plan.dest[digit]
            for g in range(world):
                sel = np.nonzero(to == g)[0]
                off = int(plan.recv_offset[s][g])
                assert sel.size == int(plan.send_counts[s][g])
                recv_k[g][off:off + sel.size] = shards_k[s][sel]
                recv_v[g][off:off + sel.size] = shards_v[s][sel]
                written[g][off:off + sel.size] += 1
    for w in written:
        assert np.all(w == 1)  # the layout tiles every receive buffer exactly once
    out_k, out_v = [], []
    for g in range(world):
        order = np.argsort(recv_k[g], kind="stable")
        out_k.append(recv_k[g][order])
        out_v.append(recv_v[g][order])
    return plan, np.concatenate(out_k), np.concatenate(out_v)


@pytest.mark.parametrize("layout", ["bucket", "dest"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["uniform", "ref31", "ent16", "dups"])
def test_plan_exchange_gives_stable_global_sort(dmod, oracle, world, kind, layout):
    n = 20011
    shards_k, shards_v = [], []
    for r in range(world):
        if kind == "uniform":
            k = oracle.mt19937_u32(10 + r, n + 13 * r)
        elif kind == "ref31":
            k = oracle.random_u32(1 + r, n, 0, 0xFFFFFFFF)
        elif kind == "ent16":
            k = oracle.mt19937_u32(10 + r, n) & np.uint32(0xFFFF)
        else:
            k = oracle.random_u32(1 + r, n, 0, 10)
        shards_k.append(k)
    base = 0
    for k in shards_k:
        shards_v.append(np.arange(base, base + k.size, dtype=np.uint32))
        base += k.size
    allk, allv = np.concatenate(shards_k), np.concatenate(shards_v)
    shift = dmod.choose_split_shift(int(allk.min()), int(allk.max()))
    plan, gk, gv = emulate(dmod, shards_k, shards_v, shift, layout)
    ek, ev = oracle.stable_sort_pairs(allk, allv)
    np.testing.assert_array_equal(gk, ek)
    np.testing.assert_array_equal(gv, ev)
    assert int(plan.send_counts.sum()) == allk.size
    np.testing.assert_array_equal(plan.send_counts.sum(axis=0), plan.recv_totals)
    if kind == "uniform" and world > 1:
        assert plan.recv_totals.max() < 1.1 * allk.size / world


def emulate_dma(dmod, shards_k, shards_v, shift, tile):
    """numpy emulation of the exchange style "dma": tile-aligned bucket-major staging arrays, ONE chunk copy per
    (source, destination), then the local sort reading its buckets as runs through DmaExchangePlan.run_table — the
    same tables the GPU path uploads (glu_radix_sort_u32kv_segmented_runs)."""
    world = len(shards_k)
    hist_all = np.stack([np.bincount((k >> shift) & 0xFF, minlength=256) for k in shards_k])
    plan = dmod.plan_dma_exchange(hist_all, tile)
    PAD = np.uint32(0xDEADBEEF)
    recv_k = [np.full(int(t) * tile, PAD, dtype=np.uint32) for t in plan.recv_tiles]
    recv_v = [np.full(int(t) * tile, PAD, dtype=np.uint32) for t in plan.recv_tiles]
    hits = [np.zeros(int(t), dtype=np.int32) for t in plan.recv_tiles]
    for s in range(world):
        digit = (shards_k[s] >> shift) & 0xFF
        stage_k = np.full(int(plan.stage_tile[s, 256]) * tile, PAD, dtype=np.uint32)
        stage_v = np.full(int(plan.stage_tile[s, 256]) * tile, PAD, dtype=np.uint32)
        for b in range(256):
            sel = np.nonzero(digit == b)[0]
            at = int(plan.stage_tile[s, b]) * tile
            stage_k[at:at + sel.size] = shards_k[s][sel]
            stage_v[at:at + sel.size] = shards_v[s][sel]
        for g in range(world):  # one contiguous chunk of whole tiles per destination
            nt = int(plan.chunk_tiles[s, g])
            src = int(plan.stage_tile[s, plan.first_bucket[g]]) * tile
            dst = int(plan.recv_base_tile[g, s]) * tile
            recv_k[g][dst:dst + nt * tile] = stage_k[src:src + nt * tile]
            recv_v[g][dst:dst + nt * tile] = stage_v[src:src + nt * tile]
            hits[g][dst // tile:dst // tile + nt] += 1
    for h in hits:
        assert np.all(h == 1)  # the chunks tile every receive array exactly once
    out_k, out_v = [], []
    for g in range(world):
        runs, seg_count, nb = plan.run_table(g)
        R = runs.shape[1] - 1
        assert R == nb * world and int(runs[2, :R].sum()) == int(plan.recv_totals[g]) == int(seg_count.sum())
        assert int(runs[0, R]) == int(plan.recv_tiles[g])
        seg_k = [[] for _ in range(nb)]
        seg_v = [[] for _ in range(nb)]
        for r in range(R):
            c, at, sg = int(runs[2, r]), int(runs[1, r]) * tile, int(runs[3, r])
            assert int(runs[0, r + 1]) - int(runs[0, r]) == -(-c // tile)
            assert int(runs[4, r]) == int(runs[0, sg * world])
            seg_k[sg].append(recv_k[g][at:at + c])
            seg_v[sg].append(recv_v[g][at:at + c])
        for sg in range(nb):
            k, v = np.concatenate(seg_k[sg]), np.concatenate(seg_v[sg])
            assert int(seg_count[sg]) == k.size
            low = k & np.uint32((1 << shift) - 1) if shift > 0 else np.zeros_like(k)
            o = np.argsort(low, kind="stable")  # the local sort only looks at the bits below the split digit
            out_k.append(k[o])
            out_v.append(v[o])
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint32)
    return plan, cat(out_k), cat(out_v)


@pytest.mark.parametrize("tile", [8, 7680])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["uniform", "ent16", "dups", "one_bucket"])
def test_plan_dma_exchange_gives_stable_global_sort(dmod, oracle, world, kind, tile):
    n = 20011
    shards_k, shards_v = [], []
    for r in range(world):
        if kind == "uniform":
            k = oracle.mt19937_u32(10 + r, n + 13 * r)
        elif kind == "ent16":
            k = oracle.mt19937_u32(10 + r, n) & np.uint32(0xFFFF)
        elif kind == "dups":
            k = oracle.random_u32(1 + r, n, 0, 10)
        else:
            k = np.full(n, 0x12345678, dtype=np.uint32)
        shards_k.append(k)
    base = 0
    for k in shards_k:
        shards_v.append(np.arange(base, base + k.size, dtype=np.uint32))
        base += k.size
    allk, allv = np.concatenate(shards_k), np.concatenate(shards_v)
    shift = dmod.choose_split_shift(int(allk.min()), int(allk.max()))
    plan, gk, gv = emulate_dma(dmod, shards_k, shards_v, shift, tile)
    ek, ev = oracle.stable_sort_pairs(allk, allv)
    np.testing.assert_array_equal(gk, ek)
    np.testing.assert_array_equal(gv, ev)
    # what the run tables promise the device code
    for g in range(world):
        runs, _, nb = plan.run_table(g)
        assert nb <= 256 and runs.shape[1] - 1 <= 256 * world
        assert int(plan.recv_tiles[g]) <= -(-int(plan.recv_totals[g]) // tile) + 256 * world


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["GLU_ROOT"])
import __graft_entry__ as entry
import oracle
glu = entry.load_package()
dmod = glu.distributed
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 30011 + 17 * rank
keys = oracle.mt19937_u32(100 + rank, n)
if os.environ.get("GLU_KIND") == "dups":
    keys = keys % np.uint32(7)
counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(counts, torch.tensor([n], dtype=torch.int64))
base = int(sum(int(c) for c in counts[:rank]))
vals = np.arange(base, base + n, dtype=np.uint32)
# min / max -> split digit (all-gather, as DistributedRadixSort(split_shift="auto") does)
mm = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(mm, torch.tensor([int(keys.min()), int(keys.max())], dtype=torch.int64))
shift = dmod.choose_split_shift(min(int(x[0]) for x in mm), max(int(x[1]) for x in mm))
# 1-2. histograms, all-gathered
digit = (keys >> shift) & 0xFF
hist = torch.from_numpy(np.bincount(digit, minlength=256).astype(np.int64))
hists = [torch.zeros(256, dtype=torch.int64) for _ in range(world)]
dist.all_gather(hists, hist)
plan = dmod.plan_exchange(torch.stack(hists).numpy())
# 4. stable local partition (the GPU path does this in glu_radix_partition_u32kv), then the planned all-to-all
order = np.argsort(digit, kind="stable")
stage_k = torch.from_numpy(keys[order].view(np.int32).copy())
stage_v = torch.from_numpy(vals[order].view(np.int32).copy())
m = int(plan.recv_totals[rank])
in_splits = [int(x) for x in plan.send_counts[rank]]
out_splits = [int(x) for x in plan.send_counts[:, rank]]
rk = torch.zeros(m, dtype=torch.int32)
rv = torch.zeros(m, dtype=torch.int32)
dist.all_to_all_single(rk, stage_k, out_splits, in_splits)
dist.all_to_all_single(rv, stage_v, out_splits, in_splits)
rk, rv = rk.numpy().view(np.uint32), rv.numpy().view(np.uint32)
o = np.argsort(rk, kind="stable")
rk, rv = rk[o], rv[o]
# gather everything on rank 0 and compare with the oracle on the concatenated input
gathered = [None] * world
dist.gather_object((keys, vals, rk, rv), gathered if rank == 0 else None, dst=0)
if rank == 0:
    allk = np.concatenate([g[0] for g in gathered]); allv = np.concatenate([g[1] for g in gathered])
    gk = np.concatenate([g[2] for g in gathered]); gv = np.concatenate([g[3] for g in gathered])
    ek, ev = oracle.stable_sort_pairs(allk, allv)
    assert np.array_equal(gk, ek) and np.array_equal(gv, ev), "distributed plan does not reproduce the stable sort"
    print("OK", world, shift, [int(x) for x in plan.recv_totals])

# ---- the same job through the "dma" exchange layout (plan_dma_exchange): tile-aligned bucket-major staging arrays, ONE
# chunk of whole tiles per (source, destination) — moved here by all_to_all_single, on the GPU by one copy-engine transfer
# per peer — and the receiver reading its buckets as runs through the run table the GPU path uploads
TILE = 64
dplan = dmod.plan_dma_exchange(torch.stack(hists).numpy(), TILE)
PAD = np.uint32(0xDEADBEEF)
stage_k = np.full(int(dplan.stage_tile[rank, 256]) * TILE, PAD, dtype=np.uint32)
stage_v = np.full(int(dplan.stage_tile[rank, 256]) * TILE, PAD, dtype=np.uint32)
for b in range(256):
    sel = np.nonzero(digit == b)[0]
    at = int(dplan.stage_tile[rank, b]) * TILE
    stage_k[at:at + sel.size] = keys[sel]
    stage_v[at:at + sel.size] = vals[sel]
send_splits = [int(dplan.chunk_tiles[rank, g]) * TILE for g in range(world)]
recv_splits = [int(dplan.chunk_tiles[s, rank]) * TILE for s in range(world)]
assert sum(send_splits) == stage_k.size and sum(recv_splits) == int(dplan.recv_tiles[rank]) * TILE
rk2 = torch.zeros(sum(recv_splits), dtype=torch.int32)
rv2 = torch.zeros(sum(recv_splits), dtype=torch.int32)
dist.all_to_all_single(rk2, torch.from_numpy(stage_k.view(np.int32).copy()), recv_splits, send_splits)
dist.all_to_all_single(rv2, torch.from_numpy(stage_v.view(np.int32).copy()), recv_splits, send_splits)
rk2, rv2 = rk2.numpy().view(np.uint32), rv2.numpy().view(np.uint32)
assert [int(x) * TILE for x in dplan.recv_base_tile[rank]] == list(np.cumsum([0] + recv_splits[:-1]))
runs, seg_count, nb = dplan.run_table(rank)
out_k, out_v = [], []
for sg in range(nb):
    ks, vs = [], []
    for r in range(sg * world, (sg + 1) * world):
        c, at = int(runs[2, r]), int(runs[1, r]) * TILE
        ks.append(rk2[at:at + c]); vs.append(rv2[at:at + c])
    k, v = np.concatenate(ks), np.concatenate(vs)
    assert k.size == int(seg_count[sg])
    low = k & np.uint32((1 << shift) - 1) if shift > 0 else np.zeros_like(k)
    o = np.argsort(low, kind="stable")   # the segmented local sort: key bits below the split digit only
    out_k.append(k[o]); out_v.append(v[o])
ok2 = np.concatenate(out_k) if out_k else np.zeros(0, np.uint32)
ov2 = np.concatenate(out_v) if out_v else np.zeros(0, np.uint32)
gathered = [None] * world
dist.gather_object((ok2, ov2), gathered if rank == 0 else None, dst=0)
if rank == 0:
    gk = np.concatenate([g[0] for g in gathered]); gv = np.concatenate([g[1] for g in gathered])
    assert np.array_equal(gk, ek) and np.array_equal(gv, ev), "dma exchange layout does not reproduce the stable sort"
    print("OK-DMA", world, [int(x) for x in dplan.recv_tiles])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("kind", ["uniform", "dups"])
def test_gloo_world2_exchange(tmp_path, kind, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GLU_ROOT=ROOT, GLU_KIND=kind, OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() % 500) + (1 if kind == "dups" else 0)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "OK 2" in r.stdout and "OK-DMA 2" in r.stdout
