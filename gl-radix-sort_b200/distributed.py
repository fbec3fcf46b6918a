"""Multi-GPU composition of the three primitives: one process per GPU, ``torch.distributed`` (NCCL) for the plumbing.

The reference is single-GPU; this module is what BASELINE.json's north_star adds on top:

* ``DistributedReduce`` / ``DistributedBlellochScan`` shard by contiguous ranges (rank r owns the r-th range).
  Each rank reduces its shard with the local kernel, the per-rank partials are all-gathered (a few bytes over
  NVLink) and combined on the device — no host synchronisation, no extra pass over the data: the scan feeds the
  rank's base into the single-pass scan as its ``init`` (``glu_scan_exclusive_init``).
* ``DistributedRadixSort`` is an MSD split followed by a local sort:
    1. every rank builds the 256-bin histogram of the split digit of its keys (``glu_radix_histogram_u32``);
    2. the histograms are all-gathered, so every rank knows ``counts[src][bucket]`` exactly;
    3. buckets are assigned to GPUs by balanced prefix (contiguous bucket ranges, ``assign_buckets``);
    4. ONE partition pass per GPU (``glu_radix_partition_by_dest_u32kv``, a onesweep pass whose "digit" is the
       destination rank of the key's bucket) scatters every pair straight into its destination GPU's receive buffer
       through NVLink peer pointers — the partition and the all-to-all are the same kernel (``exchange="p2p"``,
       CUDA IPC mappings of the peers' buffers).  A destination receives the sources' contributions one after the
       other in rank order, each in source order, which keeps the global sort stable.  ``exchange="nccl"`` is
       the two-step variant: partition into a local staging buffer, then ``all_to_all_single``;
    5. every rank sorts what it received with the local onesweep sort.  The concatenation of the ranks' outputs
       in rank order is the stable sort of the concatenation of the inputs.

The host-side planning (``choose_split_shift``, ``assign_buckets``, ``plan_exchange``) is plain numpy so that it can be
tested on CPU with the gloo backend (tests/test_distributed_cpu.py).  Everything that touches data runs in
libglu_b200.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

RADIX_BITS = 8
RADIX = 1 << RADIX_BITS


# ------------------------------------------------------------------------------------------------ host-side planning

def choose_split_shift(key_min: int, key_max: int, bits: int = RADIX_BITS) -> int:
    """Shift of the highest `bits`-wide digit in which the keys of [key_min, key_max] can differ.

    All bits above the digit are equal in every key, so bucket order == key order.  Uniform 32-bit keys give
    24 (the top byte, north_star's "top-8-bit histogram"); keys that only use their low 16 bits give 8."""
    varying = (int(key_min) ^ int(key_max)).bit_length()
    return max(0, varying - bits)


def assign_buckets(global_counts: np.ndarray, world: int) -> np.ndarray:
    """Destination rank of every bucket: contiguous, monotone bucket ranges with about total/world pairs each.

    A bucket goes to the rank in whose share of the global order its midpoint falls."""
    counts = np.asarray(global_counts, dtype=np.int64)
    total = int(counts.sum())
    if total == 0:
        return np.zeros(counts.size, dtype=np.int64)
    exclusive = np.cumsum(counts) - counts
    dest = (2 * exclusive + counts) * world // (2 * total)
    return np.minimum(dest, world - 1).astype(np.int64)


@dataclass
class ExchangePlan:
    dest: np.ndarray          # [RADIX]        destination rank of each bucket
    send_counts: np.ndarray   # [world, world] pairs rank s sends to rank g
    recv_totals: np.ndarray   # [world]        pairs each rank ends up with
    recv_offset: np.ndarray   # [world, world] where rank s's pairs start in rank g's receive buffer (source-major)
    send_offset: np.ndarray   # [world, world] where the pairs for rank g start in rank s's destination-major staging
    dst_offset: np.ndarray    # [world, RADIX] bucket-major layout: where rank s's run of bucket b starts in dest[b]'s buffer
    src_offset: np.ndarray    # [world, RADIX] where bucket b starts in rank s's own bucket-major order


def plan_exchange(hist_all: np.ndarray) -> ExchangePlan:
    """From counts[src][bucket] (the all-gathered histograms) to the complete layout of the all-to-all.

    Two receive layouts are planned, both stable (equal keys share a bucket, hence a destination, and stay in global
    input order: source rank, then source position):
      * source-major (recv_offset): rank g's buffer is [pairs from rank 0][pairs from rank 1]..., each in source order —
        what the partition-by-destination pass and the NCCL all-to-all produce;
      * bucket-major (dst_offset): g's buckets in increasing order, inside a bucket the sources in rank order — what
        the 256-way pointer-table partition produces (worlds of more than 16 ranks)."""
    counts = np.asarray(hist_all, dtype=np.int64)
    world, radix = counts.shape
    dest = assign_buckets(counts.sum(axis=0), world)
    onehot = (dest[:, None] == np.arange(world)[None, :]).astype(np.int64)  # [RADIX, world]
    send_counts = counts @ onehot                                           # [src, dest]
    recv_totals = send_counts.sum(axis=0)
    recv_offset = np.cumsum(send_counts, axis=0) - send_counts
    send_offset = np.cumsum(send_counts, axis=1) - send_counts
    # bucket b starts in dest[b]'s buffer after all earlier buckets of the same destination ...
    bucket_totals = counts.sum(axis=0)
    before = np.cumsum(bucket_totals) - bucket_totals                       # pairs in buckets < b (all destinations)
    first_of_dest = np.concatenate([[0], np.cumsum(recv_totals)[:-1]])      # pairs in destinations < dest[b]
    bucket_base = before - first_of_dest[dest]                              # dest ranges are contiguous in b
    # ... and inside the bucket the sources follow each other in rank order
    dst_offset = bucket_base[None, :] + (np.cumsum(counts, axis=0) - counts)
    src_offset = np.cumsum(counts, axis=1) - counts
    return ExchangePlan(dest, send_counts, recv_totals, recv_offset, send_offset, dst_offset, src_offset)


@dataclass
class DmaExchangePlan:
    """Layout of the exchange whose all-to-all is ONE copy-engine transfer per peer and array (exchange style "dma").

    Every rank lays its pairs out bucket-major in its staging arrays, every bucket at a tile boundary; the buckets of one
    destination are contiguous there, so what rank s sends to rank g is one chunk of whole tiles.  Rank g receives the
    chunks one after the other in source order; the run of bucket b from source s is then at a tile the plan knows, and
    the local sort reads bucket b as the runs (b, 0), (b, 1), ... (glu_radix_sort_u32kv_segmented_runs)."""
    tile: int
    counts: np.ndarray          # [world, RADIX]     the all-gathered histograms
    dest: np.ndarray            # [RADIX]            destination rank of each bucket (monotone)
    first_bucket: np.ndarray    # [world + 1]        rank g owns buckets [first_bucket[g], first_bucket[g + 1])
    stage_tile: np.ndarray      # [world, RADIX + 1] first tile of bucket b in rank s's staging arrays
    chunk_tiles: np.ndarray     # [world, world]     tiles rank s sends to rank g
    recv_base_tile: np.ndarray  # [world, world]     [g, s]: first tile of rank s's chunk in rank g's receive arrays
    recv_totals: np.ndarray     # [world]            pairs every rank receives
    recv_tiles: np.ndarray      # [world]            tiles every rank's receive arrays hold afterwards

    def phys_tile(self, g: int, b, s):
        """Tile of rank g's receive arrays where source s's run of bucket b (owned by g) starts."""
        return self.recv_base_tile[g, s] + self.stage_tile[s, b] - self.stage_tile[s, self.first_bucket[g]]

    def run_table(self, g: int):
        """(runs [5, R + 1] uint32, seg_count [RADIX] uint32, number of segments) of rank g's local sort: the runs in
        bucket order, inside a bucket in source order (include/glu_b200.h: glu_radix_sort_u32kv_segmented_runs)."""
        world = self.counts.shape[0]
        b0, b1 = int(self.first_bucket[g]), int(self.first_bucket[g + 1])
        nb = b1 - b0
        seg_count = np.zeros(RADIX, dtype=np.uint32)
        if nb == 0:
            return np.zeros((5, 1), dtype=np.uint32), seg_count, 0
        cnt = self.counts[:, b0:b1].T.reshape(-1)                       # [bucket, source]
        tiles = -(-cnt // self.tile)
        first = np.concatenate([[0], np.cumsum(tiles)])
        phys = (self.recv_base_tile[g][None, :] + (self.stage_tile[:, b0:b1] - self.stage_tile[:, b0:b0 + 1]).T).reshape(-1)
        seg = np.repeat(np.arange(nb), world)
        runs = np.zeros((5, nb * world + 1), dtype=np.uint32)
        runs[0] = first
        runs[1, :-1] = phys
        runs[2, :-1] = cnt
        runs[3, :-1] = seg
        runs[4, :-1] = first[seg * world]
        seg_count[:nb] = self.counts[:, b0:b1].sum(axis=0)
        return runs, seg_count, nb


def plan_dma_exchange(hist_all: np.ndarray, tile: int) -> DmaExchangePlan:
    counts = np.asarray(hist_all, dtype=np.int64)
    world, radix = counts.shape
    dest = assign_buckets(counts.sum(axis=0), world)
    first_bucket = np.searchsorted(dest, np.arange(world + 1), side="left").astype(np.int64)
    tiles = -(-counts // tile)
    stage_tile = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(tiles, axis=1)], axis=1)
    chunk_tiles = stage_tile[:, first_bucket[1:]] - stage_tile[:, first_bucket[:-1]]      # [src, dst]
    recv_base_tile = (np.cumsum(chunk_tiles, axis=0) - chunk_tiles).T.copy()              # [dst, src]
    onehot = (dest[:, None] == np.arange(world)[None, :]).astype(np.int64)
    recv_totals = (counts @ onehot).sum(axis=0)
    return DmaExchangePlan(int(tile), counts, dest, first_bucket, stage_tile, chunk_tiles, recv_base_tile, recv_totals,
                           chunk_tiles.sum(axis=0))


# ------------------------------------------------------------------------------------------------------ GPU side

def _glu():
    import sys

    return sys.modules[__name__.rsplit(".", 1)[0]]


def _dist():
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group('nccl'))")
    return dist


class _DeviceArray:
    """A glu_malloc'ed (cudaMalloc, hence CUDA-IPC exportable) int32 array viewed as a torch tensor."""

    def __init__(self, n_elems: int, device):
        import torch

        glu = _glu()
        self.ptr = ctypes.c_void_p()
        self.n = int(n_elems)
        glu.check(glu.lib.glu_malloc(ctypes.byref(self.ptr), 4 * max(self.n, 1)), "glu_malloc")
        self.__cuda_array_interface__ = {"shape": (max(self.n, 1),), "typestr": "<i4", "data": (self.ptr.value, False),
                                         "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)

    def free(self):
        if self.ptr:
            self.tensor = None
            _glu().lib.glu_free(self.ptr)
            self.ptr = ctypes.c_void_p()


class DistributedReduce:
    """Reduce over a buffer sharded by contiguous ranges: after the call element 0 of EVERY rank's shard holds the
    global result (glu::Reduce leaves its result in element 0, glu/Reduce.hpp:111-135)."""

    def __init__(self, data_type, operator_, group=None):
        glu = _glu()
        self._local = glu.Reduce(data_type, operator_)  # validates like the reference
        self.data_type, self.operator, self.group = self._local.data_type, self._local.operator, group
        self._esz = glu.data_type_size(self.data_type)
        self._gathered = None

    def __call__(self, buffer, count: int) -> None:
        import torch

        glu, dist = _glu(), _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if count <= 0:
            raise glu.GluError(1, "Count must be greater than zero (every rank owns a non-empty range)")
        ptr, dev = glu._ptr_and_device(buffer)
        dev = glu._device_of(dev)
        if self._gathered is None or self._gathered.device != dev:
            self._gathered = torch.empty((world + 1) * self._esz, dtype=torch.uint8, device=dev)
        gathered = self._gathered
        partial = gathered[world * self._esz:]
        need = int(glu.lib.glu_reduce_tmp_bytes(count, int(self.data_type)))
        tmp, tmp_bytes = self._local._scratch.ensure(need, dev)
        st = glu._current_stream(dev)
        glu.check(glu.lib.glu_reduce_into(ptr, count, int(self.data_type), int(self.operator), partial.data_ptr(),
                                          tmp, tmp_bytes, st), "DistributedReduce (local)")
        dist.all_gather_into_tensor(gathered[: world * self._esz], partial, group=self.group)
        glu.check(glu.lib.glu_reduce_into(gathered.data_ptr(), world, int(self.data_type), int(self.operator), ptr,
                                          tmp, tmp_bytes, st), "DistributedReduce (combine)")


class DistributedBlellochScan:
    """Exclusive prefix sum of ONE sequence sharded by contiguous ranges in rank order.  12 B of HBM traffic per
    4-byte element: a reduce pass for the rank totals (4 B), an all-gather of `world` elements, and the
    single-pass scan seeded with the rank's base (8 B)."""

    def __init__(self, data_type, group=None):
        glu = _glu()
        self._scan = glu.BlellochScan(data_type)
        self._reduce = glu.Reduce(data_type, glu.ReduceOperator_Sum)
        self.data_type, self.group = self._scan.data_type, group
        self._esz = glu.data_type_size(self.data_type)
        self._gathered = None
        self._small = glu._Scratch()

    def __call__(self, buffer, count: int) -> None:
        import torch

        glu, dist = _glu(), _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if count <= 0:
            raise glu.GluError(1, "Count must be greater than zero (every rank owns a non-empty range)")
        ptr, dev = glu._ptr_and_device(buffer)
        dev = glu._device_of(dev)
        dt, esz = int(self.data_type), self._esz
        if self._gathered is None or self._gathered.device != dev:
            self._gathered = torch.empty((world + 1) * esz, dtype=torch.uint8, device=dev)
        gathered = self._gathered
        total = gathered[world * esz:]
        st = glu._current_stream(dev)
        tmp, tmp_bytes = self._reduce._scratch.ensure(int(glu.lib.glu_reduce_tmp_bytes(count, dt)), dev)
        glu.check(glu.lib.glu_reduce_into(ptr, count, dt, int(glu.ReduceOperator_Sum), total.data_ptr(), tmp, tmp_bytes,
                                          st), "DistributedBlellochScan (rank total)")
        dist.all_gather_into_tensor(gathered[: world * esz], total, group=self.group)
        # exclusive scan of the `world` totals: element `rank` becomes this rank's base
        stmp, stmp_bytes = self._small.ensure(int(glu.lib.glu_scan_exclusive_tmp_bytes(world, 1, dt)), dev)
        glu.check(glu.lib.glu_scan_exclusive(gathered.data_ptr(), world, 1, dt, stmp, stmp_bytes, st),
                  "DistributedBlellochScan (bases)")
        need = int(glu.lib.glu_scan_exclusive_tmp_bytes(count, 1, dt))
        tmp, tmp_bytes = self._scan._scratch.ensure(need, dev)
        glu.check(glu.lib.glu_scan_exclusive_init(ptr, count, 1, dt, gathered.data_ptr() + rank * esz, tmp, tmp_bytes,
                                                  st), "DistributedBlellochScan (seeded scan)")


class DistributedRadixSort:
    """Stable sort of uint32 (key, value) pairs spread over the ranks of `group` (MSD split + local sort).

    ``sorter(keys, vals, count)`` takes this rank's `count` pairs (CUDA tensors or device pointers; not modified)
    and returns ``(sorted_keys, sorted_vals, m)``: int32-typed tensor views (uint32 bit patterns) of this rank's
    slice of the global result, valid until the next call.  Rank r's keys are <= rank r+1's."""

    def __init__(self, max_count: int, group=None, capacity_factor: float = 1.25, exchange: str = "auto",
                 split_shift: int | str = 32 - RADIX_BITS, plan: str = "auto", local: str = "auto"):
        import os

        import torch

        glu, dist = _glu(), _dist()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_count = int(max_count)
        self.capacity = int(max_count * capacity_factor) + 1024
        self.split_shift = split_shift
        # local="segmented": the exchange is BUCKET-major (one run per top-digit bucket and source, every bucket at a tile
        # boundary of the receiver) and the local sort is glu_radix_sort_u32kv_segmented over the key bits below the split
        # digit — 3 passes for 32-bit keys, north_star's "local onesweep on the remaining 24 bits".  local="full": the
        # exchange is destination-major (long NVLink runs) and the received range is sorted on all 32 bits (4 passes).
        if local == "auto":  # GLU_DIST_LOCAL chooses for callers that do not (bench.py's sweeps, the pipeline's lanes)
            local = os.environ.get("GLU_DIST_LOCAL", "auto")
        if local not in ("auto", "full", "segmented"):
            raise ValueError(local)
        self._tile = int(glu.lib.glu_radix_sort_segment_tile())
        # How the bucket-major exchange crosses NVLink.  "staged": a local MSD pass into this rank's own bucket-major
        # staging arrays (full HBM speed), then a copy kernel that moves each bucket's long run with full-line
        # stores.  "direct": the MSD pass stores straight into the peers' memory (no staging, 16 B of HBM traffic less
        # per pair, but ~30-pair runs: byte-masked NVLink packets).  "dma": the staging arrays are tile-aligned per
        # bucket, what a rank sends to a peer is ONE contiguous chunk, moved by the copy engines (cudaMemcpyAsync over
        # the peer mapping: no SM is spent on the all-to-all); the plan is made on the host from the all-gathered
        # histograms (DmaExchangePlan) and the local sort reads its buckets as runs.
        self.exchange_style = os.environ.get("GLU_DIST_EXCHANGE_STYLE", "dma")
        if self.exchange_style not in ("staged", "direct", "dma"):
            raise ValueError(self.exchange_style)
        # every bucket (dma: every run = bucket x source) may end in a partial tile
        self.capacity_tiles = -(-self.capacity // self._tile) + RADIX * (self.world if self.exchange_style == "dma" else 1)
        seg_capacity = self.capacity_tiles * self._tile
        want_seg = local in ("auto", "segmented") and exchange in ("auto", "p2p") and plan in ("auto", "device") \
            and self.world <= 16
        if max(self.capacity, seg_capacity if want_seg else 0) >= (1 << 31):
            raise glu.GluError(6, "DistributedRadixSort: per-rank capacity must stay below 2^31 pairs")
        alloc = seg_capacity if want_seg else self.capacity
        self._recv_keys = _DeviceArray(alloc, self.device)
        self._recv_vals = _DeviceArray(alloc, self.device)
        self._sorter = glu.RadixSort()
        self._alt_keys = self._alt_vals = None
        self._part_tmp = torch.empty(int(glu.lib.glu_radix_partition_u32kv_tmp_bytes(self.max_count)), dtype=torch.uint8,
                                     device=self.device)
        self._hist = torch.zeros(RADIX, dtype=torch.int32, device=self.device)
        self._hist_all = torch.zeros(self.world * RADIX, dtype=torch.int32, device=self.device)
        # [256 key pointers][256 value pointers][256-byte digit -> destination table, as 32 int64]
        self._tables = torch.zeros(2 * RADIX + RADIX // 8, dtype=torch.int64, device=self.device)
        self._tables_host = torch.zeros(2 * RADIX + RADIX // 8, dtype=torch.int64).pin_memory()
        self.by_dest = self.world <= 16
        self._minmax = None
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._peer_keys = self._peer_vals = self._peer_flags = None
        self._flags = None
        self._stage_keys = self._stage_vals = None
        # dma style: how the receiver learns that the peers' chunks have landed — "flags": every sender stores the
        # exchange's epoch into the receiver's flag array after its copies (stream-ordered), the receiver's sorting stream
        # waits for the flags (no collective; the copies may then run on a stream of their own, `_copy_stream`, set by
        # DistributedSortPipeline); "nccl": a 4-byte all-reduce after the copies
        self._dma_sync = os.environ.get("GLU_DIST_DMA_SYNC", "flags")
        if self._dma_sync not in ("flags", "nccl"):
            raise ValueError(self._dma_sync)
        self._copy_stream = None
        self._trace = None
        self._epoch = 0
        self.exchange = self._setup_exchange(exchange)
        self.local = "segmented" if (want_seg and self.exchange == "p2p") else "full"
        if local == "segmented" and self.local != "segmented":
            raise glu.GluError(1, "DistributedRadixSort: local='segmented' needs the p2p exchange, the device plan and "
                                  "at most 16 ranks")
        if self.local == "segmented":
            self._alt_keys = torch.empty(seg_capacity, dtype=torch.int32, device=self.device)
            self._alt_vals = torch.empty(seg_capacity, dtype=torch.int32, device=self.device)
            self._seg_count = torch.zeros(RADIX, dtype=torch.int32, device=self.device)
            if self.exchange_style == "dma":
                self._stage_tiles = -(-self.max_count // self._tile) + RADIX
                self._stage_k = torch.empty(self._stage_tiles * self._tile, dtype=torch.int32, device=self.device)
                self._stage_v = torch.empty(self._stage_tiles * self._tile, dtype=torch.int32, device=self.device)
                # host plan -> device: [256 key pointers][256 value pointers] of the local MSD pass, the segment counts
                # and the run table of the local sort, in one pinned block / one upload
                self._max_runs = RADIX * self.world
                words = 2 * (2 * RADIX) + RADIX + 5 * (self._max_runs + 1)   # int64 pointers as 2 words each
                self._dma_host = torch.zeros(words, dtype=torch.int32).pin_memory()
                self._dma_dev = torch.zeros(words, dtype=torch.int32, device=self.device)
                self._hist_pinned = torch.zeros(self.world * RADIX, dtype=torch.int32).pin_memory()
                self._hist_event = torch.cuda.Event()
                self._msd_event = torch.cuda.Event()
                self._sent_event = torch.cuda.Event()
                self._dma_plan = None
                self._dma_segments = 0
                self._dma_runs = 0
                self._m = 0
            elif self.exchange_style == "staged":
                self._stage_k = torch.empty(self.max_count, dtype=torch.int32, device=self.device)
                self._stage_v = torch.empty(self.max_count, dtype=torch.int32, device=self.device)
                # [256 key pointers][256 value pointers] of the local MSD pass, then [256] staging offsets and [256] copy
                # counts (uint32, as 2 x 128 int64)
                self._stage_tables = torch.zeros(2 * RADIX + RADIX, dtype=torch.int64, device=self.device)
            self._sorter._scratch.ensure(int(glu.lib.glu_radix_sort_u32kv_segmented_tmp_bytes(self.capacity_tiles)),
                                         self.device)
        else:
            self.exchange_style = "by-destination"
            self._sorter.prepare_internal_buffers(self.capacity, self.device)
        # plan="device": the exchange plan is computed by glu_radix_exchange_plan and the partition / local sort read
        # their counts from device memory — the step has no host round trip between the histogram and the sort (the
        # host only waits, after everything is enqueued, for the few bytes that tell it how much it received).
        # plan="host": histogram -> host (numpy plan_exchange) -> tables uploaded; needed by the NCCL exchange (its
        # split sizes are host arguments) and by worlds of more than 16 ranks.
        if plan not in ("auto", "device", "host"):
            raise ValueError(plan)
        can_device = self.exchange == "p2p" and self.by_dest
        if plan == "device" and not can_device:
            raise glu.GluError(1, "DistributedRadixSort: plan='device' needs the p2p exchange and at most 16 ranks")
        self.plan_mode = "device" if (plan in ("auto", "device") and can_device) else "host"
        if self.plan_mode == "device":
            peers = torch.from_numpy(np.concatenate([self._peer_keys, self._peer_vals]).astype(np.int64))
            self._peers_dev = peers.to(self.device)                      # [world] key bases, [world] value bases
            self._counts_dev = torch.zeros(2, dtype=torch.int32, device=self.device)   # pairs sent, pairs received
            self._info_dev = torch.zeros(self.world + 2, dtype=torch.int64, device=self.device)
            self._info_host = torch.zeros(self.world + 2, dtype=torch.int64).pin_memory()
            self._hist_host = torch.zeros(self.world * RADIX, dtype=torch.int32).pin_memory()
            self._plan_event = torch.cuda.Event()
        self.last_plan = None
        self._last_hist = None
        self.timing = None  # set to a dict to get per-phase wall times in ms (synchronises after every phase)
        self._closed = False

    # -- peer mappings (CUDA IPC): rank g's receive buffers mapped into this process
    def _setup_exchange(self, exchange: str) -> str:
        import torch

        glu, dist = _glu(), _dist()
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError(exchange)
        ok, err = True, ""
        if exchange in ("auto", "p2p"):
            arrays = [self._recv_keys, self._recv_vals]
            if self.exchange_style == "dma":
                # one word per source rank: the epoch of the last exchange whose chunk from that rank has landed here
                self._flags = _DeviceArray(32, self.device)
                self._flags.tensor.zero_()
                torch.cuda.synchronize(self.device)
                arrays.append(self._flags)
            try:
                mine = []
                for a in arrays:
                    h = ctypes.create_string_buffer(64)
                    glu.check(glu.lib.glu_ipc_get_handle(a.ptr, h), "glu_ipc_get_handle")
                    mine.append(h.raw)
                mine = tuple(mine)
            except glu.GluError as e:  # pragma: no cover - depends on the box
                mine, ok, err = None, False, str(e)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=self.group)
            ok = ok and all(h is not None for h in handles)
            peers = [[0] * self.world for _ in arrays]
            if ok:
                try:
                    for g, h in enumerate(handles):
                        for j, a in enumerate(arrays):
                            if g == self.rank:
                                peers[j][g] = a.ptr.value
                                continue
                            p = ctypes.c_void_p()
                            glu.check(glu.lib.glu_ipc_open_handle(h[j], ctypes.byref(p)), "glu_ipc_open_handle")
                            peers[j][g] = p.value
                except glu.GluError as e:  # pragma: no cover
                    ok, err = False, str(e)
            flags = [None] * self.world
            dist.all_gather_object(flags, ok, group=self.group)
            if all(flags):
                self._peer_keys = np.array(peers[0], dtype=np.int64)
                self._peer_vals = np.array(peers[1], dtype=np.int64)
                if len(arrays) > 2:
                    self._peer_flags = np.array(peers[2], dtype=np.int64)
                return "p2p"
            if exchange == "p2p":
                raise glu.GluError(7, f"DistributedRadixSort: CUDA IPC peer mapping failed ({err})")
        self._stage_keys = torch.empty(self.max_count, dtype=torch.int32, device=self.device)
        self._stage_vals = torch.empty(self.max_count, dtype=torch.int32, device=self.device)
        return "nccl"

    def _split_shift(self, kptr: int, count: int, st: int) -> int:
        import torch

        glu, dist = _glu(), _dist()
        if self.split_shift != "auto":
            return int(self.split_shift)
        # adaptive split digit: the highest 8 bits in which any two keys of the job differ
        if self._minmax is None:
            self._minmax = torch.zeros(2 * (self.world + 1), dtype=torch.int32, device=self.device)
            self._mm_reduce = glu.Reduce(glu.DataType_Uint, glu.ReduceOperator_Min)
        mm = self._minmax
        mine = mm[2 * self.world:]
        tmp, tmp_bytes = self._mm_reduce._scratch.ensure(int(glu.lib.glu_reduce_tmp_bytes(count, 3)), self.device)
        glu.check(glu.lib.glu_reduce_into(kptr, count, 3, int(glu.ReduceOperator_Min), mine.data_ptr(), tmp, tmp_bytes, st),
                  "DistributedRadixSort (min)")
        glu.check(glu.lib.glu_reduce_into(kptr, count, 3, int(glu.ReduceOperator_Max), mine.data_ptr() + 4, tmp, tmp_bytes,
                                          st), "DistributedRadixSort (max)")
        dist.all_gather_into_tensor(mm[: 2 * self.world], mine, group=self.group)
        host = mm[: 2 * self.world].cpu().numpy().view(np.uint32).reshape(self.world, 2)
        return choose_split_shift(int(host[:, 0].min()), int(host[:, 1].max()))

    def __call__(self, key_buffer, val_buffer, count: int):
        import torch

        glu, dist = _glu(), _dist()
        kptr, kdev = glu._ptr_and_device(key_buffer)
        vptr, _ = glu._ptr_and_device(val_buffer)
        if not kptr or not vptr:
            raise glu.GluError(1, "Invalid key / value buffer")
        if count < 1 or count > self.max_count:
            raise glu.GluError(1, f"count must be in [1, {self.max_count}]")
        st = glu._current_stream(self.device)
        timing = self.timing
        if timing is not None:
            import time

            def mark(name):
                torch.cuda.synchronize()
                now = time.perf_counter()
                timing[name] = timing.get(name, 0.0) + 1e3 * (now - mark.t)
                mark.t = now
            torch.cuda.synchronize()
            mark.t = time.perf_counter()
        else:
            def mark(name):
                pass
        shift = self._split_shift(kptr, count, st)
        if self.plan_mode == "device":
            return self._call_device_plan(kptr, vptr, count, shift, st, mark)

        # 1-2. local digit histogram -> counts[src][bucket] on every rank (this all-gather also orders this call's
        #      peer writes after every rank's previous use of its receive buffers)
        glu.check(glu.lib.glu_radix_histogram_u32(kptr, count, shift, RADIX_BITS, self._hist.data_ptr(), st),
                  "glu_radix_histogram_u32")
        dist.all_gather_into_tensor(self._hist_all, self._hist, group=self.group)
        hist_all = self._hist_all.cpu().numpy().view(np.uint32).reshape(self.world, RADIX)
        mark("histogram+allgather+d2h")

        # 3. bucket -> GPU assignment and the receive layout
        plan = plan_exchange(hist_all)
        self.last_plan = plan
        m = int(plan.recv_totals[self.rank])
        if int(plan.recv_totals.max()) > self.capacity:
            raise glu.GluError(6, f"DistributedRadixSort: a rank would receive {int(plan.recv_totals.max())} pairs, "
                                  f"capacity is {self.capacity} (raise capacity_factor or use split_shift='auto')")

        # 4. partition (+ all-to-all).  Up to 16 ranks the pass partitions by destination (a tile leaves as `world`
        #    long runs — remote stores near link speed); beyond that by bucket through a 256-entry pointer table.
        tables = self._tables_host.numpy()
        r = self.rank
        if self.by_dest:
            tables[:RADIX] = 0
            tables[RADIX:2 * RADIX] = 0
            if self.exchange == "p2p":
                tables[:self.world] = self._peer_keys + 4 * plan.recv_offset[r]
                tables[RADIX:RADIX + self.world] = self._peer_vals + 4 * plan.recv_offset[r]
            else:
                tables[:self.world] = self._stage_keys.data_ptr() + 4 * plan.send_offset[r]
                tables[RADIX:RADIX + self.world] = self._stage_vals.data_ptr() + 4 * plan.send_offset[r]
            tables[2 * RADIX:].view(np.uint8)[:] = plan.dest.astype(np.uint8)
        elif self.exchange == "p2p":
            tables[:RADIX] = self._peer_keys[plan.dest] + 4 * plan.dst_offset[r]
            tables[RADIX:2 * RADIX] = self._peer_vals[plan.dest] + 4 * plan.dst_offset[r]
        else:
            tables[:RADIX] = self._stage_keys.data_ptr() + 4 * plan.src_offset[r]
            tables[RADIX:2 * RADIX] = self._stage_vals.data_ptr() + 4 * plan.src_offset[r]
        self._tables.copy_(self._tables_host, non_blocking=True)
        mark("plan+tables")
        tptr = self._tables.data_ptr()
        if self.by_dest:
            glu.check(glu.lib.glu_radix_partition_by_dest_u32kv(kptr, vptr, count, shift, RADIX_BITS, tptr + 16 * RADIX,
                                                                tptr, tptr + 8 * RADIX, self._part_tmp.data_ptr(),
                                                                self._part_tmp.numel(), st),
                      "glu_radix_partition_by_dest_u32kv")
        else:
            glu.check(glu.lib.glu_radix_partition_u32kv(kptr, vptr, count, shift, RADIX_BITS, tptr, tptr + 8 * RADIX,
                                                        self._part_tmp.data_ptr(), self._part_tmp.numel(), st),
                      "glu_radix_partition_u32kv")
        rk, rv = self._recv_keys.tensor, self._recv_vals.tensor
        if self.exchange == "p2p":
            # device-side barrier: when this tiny all-reduce completes, every rank's partition kernel has completed,
            # i.e. every pair destined to this rank has landed in its receive buffers
            dist.all_reduce(self._token, group=self.group)
        else:
            in_splits = [int(x) for x in plan.send_counts[self.rank]]
            out_splits = [int(x) for x in plan.send_counts[:, self.rank]]
            dist.all_to_all_single(rk[:m], self._stage_keys[:count], out_splits, in_splits, group=self.group)
            dist.all_to_all_single(rv[:m], self._stage_vals[:count], out_splits, in_splits, group=self.group)

        mark("partition+exchange")
        # 5. local sort of everything this rank received
        if m > 1:
            self._sorter(rk, rv, m)
        mark("local sort")
        return rk[:m], rv[:m], m

    def _call_device_plan(self, kptr: int, vptr: int, count: int, shift: int, st: int, mark):
        """The step without a host round trip: histogram -> all-gather -> plan kernel -> partition (+ NVLink all-to-all)
        -> device barrier -> local sort, all enqueued back to back; counts travel through device memory."""
        self._enqueue_exchange(kptr, vptr, count, shift, st, mark)
        mark("partition+exchange")
        rk, rv = self._enqueue_local_sort(shift, st)
        mark("local sort")
        return self._finish(rk, rv)

    def _enqueue_exchange(self, kptr: int, vptr: int, count: int, shift: int, st: int, mark=lambda name: None):
        """histogram -> all-gather -> device plan -> partition pass whose stores are the all-to-all -> device barrier
        (enqueued on the CURRENT torch stream, whose raw handle is `st`)."""
        glu, dist = _glu(), _dist()
        world, rank = self.world, self.rank
        glu.check(glu.lib.glu_radix_histogram_u32(kptr, count, shift, RADIX_BITS, self._hist.data_ptr(), st),
                  "glu_radix_histogram_u32")
        # (this all-gather also orders this call's peer writes after every rank's previous use of its receive buffers)
        dist.all_gather_into_tensor(self._hist_all, self._hist, group=self.group)
        if self.local == "segmented" and self.exchange_style == "dma":
            return self._enqueue_exchange_dma(kptr, vptr, count, shift, st, mark)
        tptr = self._tables.data_ptr()
        pptr = self._peers_dev.data_ptr()
        cptr = self._counts_dev.data_ptr()
        if self.local == "segmented":
            glu.check(glu.lib.glu_radix_exchange_plan_buckets(self._hist_all.data_ptr(), world, rank, count,
                                                              self.capacity_tiles, pptr, pptr + 8 * world, tptr,
                                                              tptr + 8 * RADIX, self._seg_count.data_ptr(), cptr,
                                                              self._info_dev.data_ptr(), st),
                      "glu_radix_exchange_plan_buckets")
        else:
            glu.check(glu.lib.glu_radix_exchange_plan(self._hist_all.data_ptr(), world, rank, count, self.capacity, pptr,
                                                      pptr + 8 * world, tptr, tptr + 8 * RADIX, tptr + 16 * RADIX, cptr,
                                                      self._info_dev.data_ptr(), st),
                      "glu_radix_exchange_plan")
        self._info_host.copy_(self._info_dev, non_blocking=True)
        self._hist_host.copy_(self._hist_all, non_blocking=True)
        self._plan_event.record()
        mark("histogram+allgather+plan")
        if self.local == "segmented" and self.exchange_style == "staged":
            # one run per (bucket, source) at the receiver, its buckets contiguous, ready for the segmented sort:
            # local MSD pass into the staging arrays, then the long-run copy over NVLink
            sptr = self._stage_tables.data_ptr()
            my_hist = self._hist_all.data_ptr() + 4 * RADIX * rank
            glu.check(glu.lib.glu_radix_exchange_stage_tables(my_hist, self._stage_k.data_ptr(), self._stage_v.data_ptr(),
                                                              self._seg_count.data_ptr(), tptr, tptr + 8 * RADIX, sptr,
                                                              sptr + 8 * RADIX, sptr + 16 * RADIX, sptr + 20 * RADIX, st),
                      "glu_radix_exchange_stage_tables")
            glu.check(glu.lib.glu_radix_partition_u32kv_dyn(kptr, vptr, cptr, count, shift, RADIX_BITS, sptr,
                                                            sptr + 8 * RADIX, self._part_tmp.data_ptr(),
                                                            self._part_tmp.numel(), st),
                      "glu_radix_partition_u32kv_dyn (local MSD pass)")
            glu.check(glu.lib.glu_radix_exchange_copy_u32kv(self._stage_k.data_ptr(), self._stage_v.data_ptr(),
                                                            sptr + 16 * RADIX, sptr + 20 * RADIX, tptr, tptr + 8 * RADIX, 0,
                                                            st),
                      "glu_radix_exchange_copy_u32kv")
        elif self.local == "segmented":
            # the MSD pass stores straight into the peers' memory
            glu.check(glu.lib.glu_radix_partition_u32kv_dyn(kptr, vptr, cptr, count, shift, RADIX_BITS, tptr,
                                                            tptr + 8 * RADIX, self._part_tmp.data_ptr(),
                                                            self._part_tmp.numel(), st),
                      "glu_radix_partition_u32kv_dyn")
        else:
            glu.check(glu.lib.glu_radix_partition_by_dest_u32kv_dyn(kptr, vptr, cptr, count, shift, RADIX_BITS,
                                                                    tptr + 16 * RADIX, tptr, tptr + 8 * RADIX,
                                                                    self._part_tmp.data_ptr(), self._part_tmp.numel(), st),
                      "glu_radix_partition_by_dest_u32kv_dyn")
        # device-side barrier: when this tiny all-reduce completes, every rank's partition kernel has completed
        dist.all_reduce(self._token, group=self.group)

    def _enqueue_exchange_dma(self, kptr: int, vptr: int, count: int, shift: int, st: int, mark):
        """Exchange style "dma" (after the histogram all-gather): the host reads the histograms, plans, uploads one
        small block of tables; a local MSD pass brings the pairs into tile-aligned bucket-major order (the buckets this
        rank keeps go straight to their place in its own receive arrays); one copy-engine transfer per peer and array
        moves the chunks; a device-side barrier.  The only host wait of the step is the one for the 8 KB of histograms."""
        import torch

        glu, dist = _glu(), _dist()
        world, rank, tile = self.world, self.rank, self._tile
        self._hist_pinned.copy_(self._hist_all, non_blocking=True)
        self._hist_event.record()
        tr = getattr(self, "_trace", None)
        if tr is not None:
            import time

            tr["hist"] = torch.cuda.Event(enable_timing=True)
            tr["hist"].record()
        self._hist_event.synchronize()
        if tr is not None:
            tr["host_hist_seen"] = time.perf_counter()
        hist_all = self._hist_pinned.numpy().view(np.uint32).reshape(world, RADIX)
        plan = plan_dma_exchange(hist_all, tile)
        self._dma_plan = plan
        self._last_hist = hist_all.copy()
        self.last_plan = None
        # every rank sees the same plan, so every rank raises (or none does)
        if int(plan.recv_tiles.max()) > self.capacity_tiles or int(plan.recv_totals.max()) > self.capacity:
            raise glu.GluError(6, f"DistributedRadixSort: a rank would receive {int(plan.recv_totals.max())} pairs, "
                                  f"capacity is {self.capacity} (raise capacity_factor or use split_shift='auto')")
        runs, seg_count, nb = plan.run_table(rank)
        self._m = int(plan.recv_totals[rank])
        self._dma_segments, self._dma_runs = nb, runs.shape[1] - 1
        host = self._dma_host.numpy()
        ptrs = host[: 4 * RADIX].view(np.int64)                       # [256 key pointers][256 value pointers]
        mine = plan.dest == rank
        b = np.arange(RADIX)
        own_tile = np.where(mine, plan.phys_tile(rank, np.where(mine, b, plan.first_bucket[rank]), rank), 0)
        stage_tile = plan.stage_tile[rank, :RADIX]
        ptrs[:RADIX] = np.where(mine, self._recv_keys.ptr.value + 4 * tile * own_tile,
                                self._stage_k.data_ptr() + 4 * tile * stage_tile)
        ptrs[RADIX:] = np.where(mine, self._recv_vals.ptr.value + 4 * tile * own_tile,
                                self._stage_v.data_ptr() + 4 * tile * stage_tile)
        host[4 * RADIX: 5 * RADIX] = seg_count.view(np.int32)
        r0 = 5 * RADIX
        host[r0: r0 + runs.size] = runs.reshape(-1).view(np.int32)
        self._dma_dev.copy_(self._dma_host, non_blocking=True)
        if tr is not None:
            tr["host_plan_done"] = time.perf_counter()
            tr["plan"] = torch.cuda.Event(enable_timing=True)
            tr["plan"].record()
        mark("histogram+allgather+plan")
        dptr = self._dma_dev.data_ptr()
        glu.check(glu.lib.glu_radix_partition_u32kv(kptr, vptr, count, shift, RADIX_BITS, dptr, dptr + 8 * RADIX,
                                                    self._part_tmp.data_ptr(), self._part_tmp.numel(), st),
                  "glu_radix_partition_u32kv (local MSD pass)")
        # the copies: on the exchange stream itself, or (pipeline) on the copy stream once the MSD pass is done, so that
        # the exchange stream is free for the next job's histogram / plan / MSD pass while the copy engines work
        cst = st
        if tr is not None:
            tr["msd"] = torch.cuda.Event(enable_timing=True)
            tr["msd"].record()
        if self._copy_stream is not None:
            self._msd_event.record()
            self._copy_stream.wait_event(self._msd_event)
            cst = self._copy_stream.cuda_stream
        for i in range(1, world):
            g = (rank + i) % world
            nt = int(plan.chunk_tiles[rank, g])
            if nt == 0:
                continue
            src = 4 * tile * int(plan.stage_tile[rank, plan.first_bucket[g]])
            dst = 4 * tile * int(plan.recv_base_tile[g, rank])
            glu.check(glu.lib.glu_memcpy_d2d(int(self._peer_keys[g]) + dst, self._stage_k.data_ptr() + src, 4 * tile * nt, cst),
                      "DistributedRadixSort (peer copy, keys)")
            glu.check(glu.lib.glu_memcpy_d2d(int(self._peer_vals[g]) + dst, self._stage_v.data_ptr() + src, 4 * tile * nt, cst),
                      "DistributedRadixSort (peer copy, values)")
        self._epoch += 1
        if self._dma_sync == "flags" and world > 1:
            addrs = (ctypes.c_uint64 * world)(*[0 if g == rank else int(self._peer_flags[g]) + 4 * rank
                                                for g in range(world)])
            glu.check(glu.lib.glu_signal_peers_u32(addrs, world, self._epoch & 0xFFFFFFFF, cst), "glu_signal_peers_u32")
        if self._copy_stream is not None:
            self._sent_event.record(self._copy_stream)
        if self._dma_sync == "nccl":
            if self._copy_stream is not None:
                torch.cuda.current_stream(self.device).wait_event(self._sent_event)
            # device-side barrier: when this tiny all-reduce completes, every rank's copies have completed
            dist.all_reduce(self._token, group=self.group)

    def _enqueue_local_sort(self, shift: int, st: int):
        """The local sort of what this rank received, on stream `st`; returns the arrays that will hold the result."""
        rk, rv = self._recv_keys.tensor, self._recv_vals.tensor
        if self.local == "segmented" and self.exchange_style == "dma":
            if self._dma_segments == 0:  # this rank owns no bucket (heavily skewed keys): it receives nothing
                return rk, rv
            if self._dma_sync == "flags" and self.world > 1:
                glu = _glu()
                glu.check(glu.lib.glu_stream_wait_flags_u32(self._flags.ptr, self.world, self.rank,
                                                            self._epoch & 0xFFFFFFFF, st), "glu_stream_wait_flags_u32")
            end_bit = shift if shift > 0 else RADIX_BITS
            dptr = self._dma_dev.data_ptr()
            in_b = self._sorter.sort_segmented(rk, rv, self._alt_keys, self._alt_vals, dptr + 4 * 4 * RADIX,
                                               self._dma_segments, self.capacity_tiles, 0, end_bit, stream=st,
                                               runs_buffer=dptr + 4 * 5 * RADIX, num_runs=self._dma_runs)
            return (self._alt_keys, self._alt_vals) if in_b else (rk, rv)
        if self.local == "segmented":
            # key bits [0, shift) are what is left to sort inside a bucket.  shift == 0 (adaptive split digit at the
            # bottom of the key): the keys of a bucket are all equal — one pass over the digit itself is a stable no-op
            # that still brings the tile-aligned buckets into the compact output layout
            end_bit = shift if shift > 0 else RADIX_BITS
            in_b = self._sorter.sort_segmented(rk, rv, self._alt_keys, self._alt_vals, self._seg_count, RADIX,
                                               self.capacity_tiles, 0, end_bit, stream=st)
            return (self._alt_keys, self._alt_vals) if in_b else (rk, rv)
        self._sorter.sort_device_count(rk, rv, self._counts_dev.data_ptr() + 4, self.capacity, stream=st)
        return rk, rv

    def _finish(self, rk, rv):
        """Host side of a step, after everything is enqueued: how much did this rank receive, did anybody overflow."""
        m = self._result_count()
        return rk[:m], rv[:m], m

    def _result_count(self) -> int:
        """Pairs this rank received in the step enqueued last; raises if any rank overflowed."""
        glu = _glu()
        world = self.world
        if self.local == "segmented" and self.exchange_style == "dma":
            return self._m  # the host made the plan (and raised there on overflow)
        # only now does the host look at the plan: the GPU is busy with the partition and the sort meanwhile
        self._plan_event.synchronize()
        info = self._info_host.numpy()
        self.last_plan = None
        self._last_hist = self._hist_host.numpy().view(np.uint32).reshape(world, RADIX).copy()
        if int(info[world + 1]) != 0:
            raise glu.GluError(6, f"DistributedRadixSort: a rank would receive {int(info[:world].max())} pairs, "
                                  f"capacity is {self.capacity} (raise capacity_factor or use split_shift='auto')")
        return int(info[world])

    def close(self, collective: bool = True) -> None:
        """Release the receive arrays and the CUDA IPC mappings of the peers' arrays.  Collective by default: a barrier
        first, so that no rank unmaps or frees memory a peer's partition pass may still be writing to."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        glu = _glu()
        try:
            import torch

            torch.cuda.synchronize(self.device)
            if collective:
                _dist().barrier(group=self.group)
        except Exception:  # pragma: no cover - interpreter shutdown / process group already destroyed
            pass
        if self._peer_keys is not None:
            for g in range(self.world):
                if g != self.rank:
                    glu.lib.glu_ipc_close_handle(ctypes.c_void_p(int(self._peer_keys[g])))
                    glu.lib.glu_ipc_close_handle(ctypes.c_void_p(int(self._peer_vals[g])))
                    if self._peer_flags is not None:
                        glu.lib.glu_ipc_close_handle(ctypes.c_void_p(int(self._peer_flags[g])))
            self._peer_keys = self._peer_vals = self._peer_flags = None
        if self._flags is not None:
            self._flags.free()
            self._flags = None
        self._recv_keys.free()
        self._recv_vals.free()
        self._sorter = None
        self._part_tmp = None
        self._alt_keys = self._alt_vals = None
        self._stage_k = self._stage_v = None

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:  # pragma: no cover
            pass

    def plan_of_last_call(self) -> ExchangePlan:
        """The exchange plan of the last call (host restatement of what the device computed, for inspection/tests)."""
        if self.last_plan is None and self._last_hist is not None:
            self.last_plan = plan_exchange(self._last_hist.copy())
        return self.last_plan


class DistributedSortPipeline:
    """A stream of independent distributed sorts with the exchange of job k+1 overlapping the local sort of job k —
    what `bench.py` times at N > 1 (world-2 and world-8 parity: tests/test_multigpu_gpu.py; every bench step's last
    output is verified on the device).

    Why: a step is the split-digit histogram, the plan, a local MSD pass, the NVLink all-to-all and a 3-pass local
    sort.  The all-to-all (2.4 ms of copy-engine time at 8 GPUs and 2^28 pairs each, no SM involved with the "dma"
    exchange style) and the host's plan hide under the previous job's local sort when consecutive jobs run on two
    streams and two lanes — each lane a complete `DistributedRadixSort` with its own receive / staging arrays.

        pipe = DistributedSortPipeline(max_count)
        t0 = pipe.submit(keys0, vals0, n)        # enqueues, returns a ticket
        t1 = pipe.submit(keys1, vals1, n)        # its exchange runs under job 0's local sort
        sk, sv, m = pipe.result(t0)              # valid until `num_lanes` later submits reuse the lane

    Ordering.  Job k uses lane k % num_lanes.  Stream X carries histogram -> all-gather -> plan -> MSD pass -> peer
    copies -> all-reduce barrier, stream S the local sort.  Before job k's all-gather, X waits for this rank's local sort
    of job k - num_lanes (same lane) and for everything the caller's stream had enqueued at submit time (its consumption
    of that job's result); the all-gather completes only when every rank has got that far, so nobody's receive arrays
    of the lane are overwritten while still in use.  All collectives are issued on X, in the same order on every rank.
    With the "dma" exchange style `submit` blocks the host until the job's histograms are all-gathered (the plan is made
    on the host); the GPU keeps working on the previous jobs meanwhile."""

    def __init__(self, max_count: int, group=None, capacity_factor: float = 1.25, split_shift: int = 32 - RADIX_BITS,
                 lanes: int | None = None):
        import os

        import torch

        glu = _glu()
        # Lanes = complete sets of receive / staging arrays and scratch.  Job k uses lane k % lanes; its exchange may
        # start as soon as the local sort of job k - lanes has finished.  Two lanes are enough: the sorting stream is the
        # longer of the two chains, so the exchange of job k never has to start before the sort of job k - 2 is over
        # (measured at 2 GPUs, 2^28 pairs each: 5.36 ms per step with two lanes, 5.41 ms with three — and a lane is
        # ~32 GB at 2^30 pairs per GPU; gpurun_out/r02j).
        if lanes is None:
            lanes = int(os.environ.get("GLU_PIPE_LANES", "2"))
        if lanes < 2:
            raise ValueError("DistributedSortPipeline needs at least two lanes")
        self.num_lanes = lanes
        self.lanes = [DistributedRadixSort(max_count, group=group, capacity_factor=capacity_factor, exchange="p2p",
                                           split_shift=int(split_shift), plan="device") for _ in range(lanes)]
        self.device = self.lanes[0].device
        # GLU_PIPE_PRIORITY=x / s gives the exchange / the sorting stream the higher CUDA stream priority (its CTAs are
        # dispatched first whenever both streams have work pending); default: equal priorities
        prio = os.environ.get("GLU_PIPE_PRIORITY", "x")
        self.stream_x = torch.cuda.Stream(device=self.device, priority=-1 if prio == "x" else 0)
        self.stream_s = torch.cuda.Stream(device=self.device, priority=-1 if prio == "s" else 0)
        # "dma" exchange with flag signalling: the peer copies of job k run on a stream of their own (copy engines only),
        # so the exchange stream goes on with job k + 1's histogram / plan / MSD pass meanwhile.  GLU_PIPE_COPY_STREAM=0:
        # the copies stay on the exchange stream.
        self.stream_d = None
        lane0 = self.lanes[0]
        if (lane0.local == "segmented" and lane0.exchange_style == "dma" and lane0._dma_sync == "flags"
                and os.environ.get("GLU_PIPE_COPY_STREAM", "1") != "0"):
            self.stream_d = torch.cuda.Stream(device=self.device, priority=-1)
            for lane in self.lanes:
                lane._copy_stream = self.stream_d
        # GLU_PIPE_TRACE=1: timing events around the phases of every job (tools/pipeline_timeline.py)
        self.trace = [] if os.environ.get("GLU_PIPE_TRACE", "0") == "1" else None
        self._exchanged = [torch.cuda.Event() for _ in range(lanes)]
        self._sorted = [torch.cuda.Event() for _ in range(lanes)]
        self._submitted = 0
        self._results = [None] * lanes
        self._glu = glu

    def submit(self, key_buffer, val_buffer, count: int) -> int:
        import torch

        glu, dist = self._glu, _dist()
        k = self._submitted
        L = self.num_lanes
        lane = self.lanes[k % L]
        kptr, _ = glu._ptr_and_device(key_buffer)
        vptr, _ = glu._ptr_and_device(val_buffer)
        if not kptr or not vptr:
            raise glu.GluError(1, "Invalid key / value buffer")
        if count < 1 or count > lane.max_count:
            raise glu.GluError(1, f"count must be in [1, {lane.max_count}]")
        shift = int(lane.split_shift)
        tr = None
        if self.trace is not None:
            import time

            def ev(stream):
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                return e
            tr = {"job": k, "host_submit": time.perf_counter()}
            lane._trace = tr
            self.trace.append(tr)
        self.stream_x.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream_x):
            if k >= L:
                self.stream_x.wait_event(self._sorted[k % L])
                if self.stream_d is not None:  # the lane's staging arrays: its previous copies have left
                    self.stream_x.wait_event(lane._sent_event)
            if tr is not None:
                tr["x_begin"] = ev(self.stream_x)
            lane._enqueue_exchange(kptr, vptr, count, shift, self.stream_x.cuda_stream)
            self._exchanged[k % L].record(self.stream_x)
            if tr is not None:
                tr["x_end"] = ev(self.stream_x)
                if self.stream_d is not None:
                    tr["copies_end"] = ev(self.stream_d)
        with torch.cuda.stream(self.stream_s):
            self.stream_s.wait_event(self._exchanged[k % L])
            if tr is not None:
                tr["s_begin"] = ev(self.stream_s)
            self._results[k % L] = lane._enqueue_local_sort(shift, self.stream_s.cuda_stream)
            self._sorted[k % L].record(self.stream_s)
            if tr is not None:
                tr["s_end"] = ev(self.stream_s)
                tr["host_return"] = time.perf_counter()
                lane._trace = None
        self._submitted = k + 1
        return k

    def result(self, ticket: int):
        """(sorted_keys, sorted_vals, m) of job `ticket`; call before `num_lanes` later submits reuse its lane.  The
        returned views are ready on the CALLER's current stream (it is made to wait for the job's local sort)."""
        import torch

        glu = self._glu
        L = self.num_lanes
        if not (self._submitted - L <= ticket < self._submitted) or ticket < 0:
            raise glu.GluError(1, "DistributedSortPipeline.result: the job's lane has been reused (or never submitted)")
        lane = self.lanes[ticket % L]
        torch.cuda.current_stream(self.device).wait_event(self._sorted[ticket % L])
        m = lane._result_count()
        rk, rv = self._results[ticket % L]
        return rk[:m], rv[:m], m

    def flush(self) -> None:
        """Make the caller's current stream wait for everything submitted so far (e.g. before recording a timing event)."""
        import torch

        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream_x)
        cur.wait_stream(self.stream_s)
        if self.stream_d is not None:
            cur.wait_stream(self.stream_d)

    def close(self) -> None:
        """Collective: releases both lanes (receive arrays, peer mappings)."""
        import torch

        torch.cuda.synchronize(self.device)
        for lane in self.lanes:
            lane.close()
