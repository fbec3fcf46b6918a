"""GPU parity tests of glu_radix_sort_u32kv_segmented — many independent stable sorts by a key-bit range in one set of
launches (the "local onesweep on the remaining 24 bits" of the multi-GPU sort, BASELINE.json north_star).  The oracle
is std::stable_sort of every segment on its own, comparing the key bits that take part (oracle.stable_sort_ex)."""
import numpy as np
import pytest

from conftest import to_device, to_host

pytestmark = pytest.mark.gpu


def run_case(glu, dev, oracle, counts, begin_bit, end_bit, seed=1, key_transform=None):
    import torch

    tile = int(glu.lib.glu_radix_sort_segment_tile())
    counts = np.asarray(counts, dtype=np.int64)
    tiles = (counts + tile - 1) // tile
    first = np.cumsum(tiles) - tiles
    max_tiles = int(tiles.sum()) + 3  # some head room: the scratch may be larger than what the segments need
    total = int(counts.sum())
    keys = oracle.mt19937_u32(seed, max(total, 1))[:total]
    if key_transform is not None:
        keys = key_transform(keys)
    vals = (np.arange(total, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(seed)  # arbitrary payload
    a_keys = np.full(max_tiles * tile, 0x0BADF00D, dtype=np.uint32)  # what lies in the padding must never matter
    a_vals = np.full(max_tiles * tile, 0xDEADDEAD, dtype=np.uint32)
    want_k, want_v = np.empty(total, np.uint32), np.empty(total, np.uint32)
    pos = 0
    for s, c in enumerate(counts):
        c = int(c)
        seg_k, seg_v = keys[pos:pos + c], vals[pos:pos + c]
        a_keys[first[s] * tile:first[s] * tile + c] = seg_k
        a_vals[first[s] * tile:first[s] * tile + c] = seg_v
        if c:
            ek, ev = oracle.stable_sort_ex(seg_k, seg_v, begin_bit, end_bit)
            want_k[pos:pos + c], want_v[pos:pos + c] = ek, ev
        pos += c
    dka, dva = to_device(a_keys, dev), to_device(a_vals, dev)
    dkb = torch.full((max_tiles * tile,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    dvb = torch.full((max_tiles * tile,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    dcount = to_device(counts.astype(np.uint32), dev)
    in_b = glu.RadixSort().sort_segmented(dka, dva, dkb, dvb, dcount, counts.size, max_tiles, begin_bit, end_bit)
    torch.cuda.synchronize()
    assert in_b == (((end_bit - begin_bit + 7) // 8) % 2 == 1)
    gk = to_host(dkb if in_b else dka, np.uint32)[:total]
    gv = to_host(dvb if in_b else dva, np.uint32)[:total]
    np.testing.assert_array_equal(gk, want_k)
    np.testing.assert_array_equal(gv, want_v)


@pytest.mark.parametrize("bits", [(0, 24), (0, 8), (0, 16), (8, 32), (0, 32), (3, 21)])
@pytest.mark.parametrize("shape", ["one_small", "empties", "exact_tile", "ragged", "many", "one_big"])
def test_segmented_sort_matches_per_segment_stable_sort(glu, cuda_device, oracle, bits, shape):
    tile = int(glu.lib.glu_radix_sort_segment_tile())
    rng = np.random.default_rng(7)
    counts = {
        "one_small": [5],
        "empties": [0, 7, 0, 0, 1, 0],
        "exact_tile": [tile, tile, 2 * tile],
        "ragged": [tile + 1, 1, 0, 3 * tile - 1, 12345, 2, tile - 1],
        "many": list(rng.integers(0, 3 * tile, size=256)),
        "one_big": [3, 50 * tile + 17, 0, 9],
    }[shape]
    run_case(glu, cuda_device, oracle, counts, bits[0], bits[1], seed=3)


def test_segmented_sort_heavy_duplicates_and_skew(glu, cuda_device, oracle):
    tile = int(glu.lib.glu_radix_sort_segment_tile())
    counts = [4 * tile + 5, 100_001, 3]
    run_case(glu, cuda_device, oracle, counts, 0, 24, seed=5, key_transform=lambda k: k % np.uint32(7))
    run_case(glu, cuda_device, oracle, counts, 0, 24, seed=6, key_transform=lambda k: np.full_like(k, 0x00ABCDEF))
    run_case(glu, cuda_device, oracle, counts, 0, 24, seed=7, key_transform=lambda k: k & np.uint32(0xFF0000FF))


def run_runs_case(glu, dev, oracle, run_counts, begin_bit, end_bit, seed=1, key_transform=None):
    """glu_radix_sort_u32kv_segmented_runs: segment s is the concatenation of the runs run_counts[s][0], [1], ...; the
    runs lie at shuffled tile positions of the A arrays.  Oracle: std::stable_sort of every segment on its own."""
    import torch

    tile = int(glu.lib.glu_radix_sort_segment_tile())
    rng = np.random.default_rng(seed)
    flat = np.array([c for seg in run_counts for c in seg], dtype=np.int64)
    seg_of = np.array([s for s, seg in enumerate(run_counts) for _ in seg], dtype=np.int64)
    run_tiles = (flat + tile - 1) // tile
    first = np.concatenate([[0], np.cumsum(run_tiles)])
    # physical placement: the runs in a random order, with random gaps of whole tiles between them
    order = rng.permutation(flat.size)
    phys = np.zeros(flat.size, dtype=np.int64)
    at = int(rng.integers(0, 3))
    for r in order:
        phys[r] = at
        at += int(run_tiles[r]) + int(rng.integers(0, 2))
    seg_counts = np.array([sum(seg) for seg in run_counts], dtype=np.int64)
    std_tiles = int(((seg_counts + tile - 1) // tile).sum())
    max_tiles = max(at, int(first[-1]), std_tiles) + 2
    total = int(flat.sum())
    keys = oracle.mt19937_u32(seed, max(total, 1))[:total]
    if key_transform is not None:
        keys = key_transform(keys)
    vals = (np.arange(total, dtype=np.uint32) * np.uint32(2654435761)) ^ np.uint32(seed)
    a_keys = np.full(max_tiles * tile, 0x0BADF00D, dtype=np.uint32)
    a_vals = np.full(max_tiles * tile, 0xDEADDEAD, dtype=np.uint32)
    want_k, want_v = np.empty(total, np.uint32), np.empty(total, np.uint32)
    pos = 0
    r = 0
    for s, seg in enumerate(run_counts):
        seg_begin = pos
        for c in seg:
            a_keys[phys[r] * tile:phys[r] * tile + c] = keys[pos:pos + c]
            a_vals[phys[r] * tile:phys[r] * tile + c] = vals[pos:pos + c]
            pos += c
            r += 1
        if pos > seg_begin:
            ek, ev = oracle.stable_sort_ex(keys[seg_begin:pos], vals[seg_begin:pos], begin_bit, end_bit)
            want_k[seg_begin:pos], want_v[seg_begin:pos] = ek, ev
    runs = np.zeros((5, flat.size + 1), dtype=np.uint32)
    runs[0] = first
    runs[1, :-1] = phys
    runs[2, :-1] = flat
    runs[3, :-1] = seg_of
    seg_first_run = np.concatenate([[0], np.cumsum([len(seg) for seg in run_counts])])[:-1]
    runs[4, :-1] = first[seg_first_run[seg_of]]
    dka, dva = to_device(a_keys, dev), to_device(a_vals, dev)
    dkb = torch.full((max_tiles * tile,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    dvb = torch.full((max_tiles * tile,), 0x5A5A5A5A, dtype=torch.int32, device=dev)
    dcount = to_device(seg_counts.astype(np.uint32), dev)
    druns = to_device(runs.reshape(-1), dev)
    in_b = glu.RadixSort().sort_segmented(dka, dva, dkb, dvb, dcount, len(run_counts), max_tiles, begin_bit, end_bit,
                                          runs_buffer=druns, num_runs=flat.size)
    torch.cuda.synchronize()
    gk = to_host(dkb if in_b else dka, np.uint32)[:total]
    gv = to_host(dvb if in_b else dva, np.uint32)[:total]
    np.testing.assert_array_equal(gk, want_k)
    np.testing.assert_array_equal(gv, want_v)


@pytest.mark.parametrize("bits", [(0, 24), (0, 8), (0, 16), (5, 32)])
@pytest.mark.parametrize("shape", ["one_run", "empties", "exact_tiles", "ragged", "many"])
def test_segmented_sort_of_runs(glu, cuda_device, oracle, bits, shape):
    tile = int(glu.lib.glu_radix_sort_segment_tile())
    rng = np.random.default_rng(11)
    run_counts = {
        "one_run": [[5]],
        "empties": [[0, 7, 0], [0, 0], [1, 0, 0, 3], [0]],
        "exact_tiles": [[tile, 2 * tile], [tile], [tile, tile, tile]],
        "ragged": [[tile + 1, 1, 0, 3 * tile - 1], [12345, 2], [tile - 1, tile + 7, 5 * tile + 3, 9]],
        "many": [list(rng.integers(0, 2 * tile, size=8)) for _ in range(32)],   # 32 buckets x 8 sources
    }[shape]
    run_runs_case(glu, cuda_device, oracle, run_counts, bits[0], bits[1], seed=3)


def test_segmented_sort_of_runs_duplicates(glu, cuda_device, oracle):
    tile = int(glu.lib.glu_radix_sort_segment_tile())
    run_counts = [[2 * tile + 5, 100_001, 3], [7, 0, tile]]
    run_runs_case(glu, cuda_device, oracle, run_counts, 0, 24, seed=5, key_transform=lambda k: k % np.uint32(7))
    run_runs_case(glu, cuda_device, oracle, run_counts, 0, 24, seed=6, key_transform=lambda k: np.full_like(k, 0x00ABCDEF))


def test_segmented_sort_argument_checks(glu, cuda_device):
    import torch

    tile = int(glu.lib.glu_radix_sort_segment_tile())
    buf = torch.zeros(4 * tile, dtype=torch.int32, device=cuda_device)
    cnt = torch.zeros(4, dtype=torch.int32, device=cuda_device)
    rs = glu.RadixSort()
    with pytest.raises(glu.GluError):
        rs.sort_segmented(buf, buf, buf, buf, cnt, 257, 4)           # too many segments
    with pytest.raises(glu.GluError):
        rs.sort_segmented(buf, buf, buf, buf, cnt, 4, 4, 8, 8)        # no key bit takes part
    with pytest.raises(glu.GluError):
        rs.sort_segmented(buf[1:], buf, buf, buf, cnt, 4, 3)          # misaligned array (bulk copies need 16 bytes)


def test_segmented_sort_other_kernel_form(cuda_device):
    """The same cases through the other form of the digit pass (GLU_SEG_KERNEL, read once per process: 0 = one tile per
    CTA, 1 = persistent ring kernel): both must agree with the oracle whatever the default is."""
    import os
    import subprocess
    import sys

    from conftest import ROOT
    if os.environ.get("GLU_SEG_TEST_CHILD"):
        pytest.skip("already the child run")
    other = "0" if os.environ.get("GLU_SEG_KERNEL", "0") == "1" else "1"
    env = dict(os.environ, GLU_SEG_KERNEL=other, GLU_SEG_TEST_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_sort_segmented_gpu.py"), "-m", "gpu",
                        "-x", "-q", "-k", "matches_per_segment or heavy_duplicates"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
