// glu_onesweep_ring.cuh — the PERSISTENT form of the onesweep digit pass (included by glu_radix_sort.cu, inside its
// anonymous namespace: it shares match_digit(), chain_cta() and the look-back word format with onesweep_kernel).
//
// Why a second form.  onesweep_kernel is one tile per CTA: every CTA starts by waiting for its own bulk copies, and
// round 1's stall samples put ~22 % of all warp time on that mbarrier (profiles/r01_onesweep_source_hotspots.md) — the
// shared-memory and register budget of three resident CTAs per SM is held by warps that wait for DRAM.  Here a CTA
// stays resident and walks over tiles it draws from an atomic ticket, with a two-deep ring of KEY staging buffers:
//
//     iteration k (tile k's keys in keys[k & 1], its digit counts already in warp_hist, its keys already in registers)
//       1. digit threads: scan of the tile's digit counts -> per-warp slot offsets
//       2. ranking warps: ballot match, keys straight to their tile-sorted slot (in place); each warp clears its
//          own counter row when it is done with it
//       3. EARLY COUNTS of tile k + 1: its keys (bulk copy issued one iteration ago into the other ring slot) go to
//          registers and into the warp's counters, and its count row is PUBLISHED — an iteration before tile k + 1 is
//          processed and BEFORE tile k's look-back: a tile's counts never wait for anything but its own bulk copy
//       4. values (bulk copy issued one iteration ago) -> registers; digit threads read ONE prefix row -> gbase;
//          values -> tile-sorted slot (in place)
//       5. tile k leaves: consecutive threads, consecutive addresses inside every digit run
//       6. one thread refills: values of tile k + 1, keys of tile k + 2 (ticket drawn an iteration earlier), L2
//          prefetch of the values of tile k + 2
//
// so a tile's copies always have most of an iteration (~10 us) to land and no warp ever waits for DRAM.  Tickets are
// drawn in increasing order and every wait of tile t (prefix row t - 1) depends only on count rows of tiles < t,
// each held by a live CTA that reaches them without waiting on anything but smaller tiles: progress by induction on
// the tile index, whatever the CTA dispatch order — two concurrent sorts on one GPU cannot deadlock each other.
// The chain CTAs (chain_cta) are the first blocks of the grid exactly as in onesweep_kernel.

template<int THREADS, int IPT, bool HAS_VALS = true> struct RingSmem
{
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * IPT;
    alignas(128) uint32_t keys[2][TILE];              // ring of TMA destinations (input order), then tile-sorted keys
    alignas(128) uint32_t vals[HAS_VALS ? TILE : 32]; // TMA destination (input order), then tile-sorted values
    alignas(16) uint32_t warp_hist[WARPS][k_radix];   // per-warp digit counts, then running slot offsets
    uint32_t gbase[k_radix];                          // global index of tile-sorted slot 0, per digit
    uint32_t tile_start[k_radix];                     // first tile-sorted slot of each digit
    uint32_t scan[8];
    alignas(8) uint64_t bar_keys[2];                  // mbarriers completed by the bulk copies
    alignas(8) uint64_t bar_vals;
    uint32_t tile_of[2];                              // tile whose keys are (or will be) in keys[b]
    uint2 info_of[2];                                 // SEG: {valid | segment << 24, first tile of the segment} of that tile
};

// RATOM: the ranking loop takes the running slot offset with ONE returning shared atomic per digit group (its
// lowest lane) and a shuffle, instead of every peer reading and re-writing the counter between two warp barriers.
//
// SEG (glu_radix_sort_seg.cuh): many independent segments in one launch.  Tile t is always elements [t * TILE, (t + 1) *
// TILE) of the input (segments start at tile boundaries), tile_info[t] = {valid | segment << 24, first tile of the
// segment}; slots past `valid` are padding.  digit_offset is [segment][256] and includes the segment's output base; the
// chain CTAs keep one running prefix over all tiles and a tile subtracts the prefix row in front of its segment.
// *d_n is then the number of TILES (device-resident).
template<int THREADS, int IPT, int MIN_BLOCKS, int MODE, int RATOM = 0, int FLAVOR = 0, bool SEG = false>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    onesweep_ring_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                         uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, uint32_t shift,
                         uint32_t mask, const uint32_t* __restrict__ digit_offset, uint32_t* lookback, uint32_t* prefix,
                         uint32_t* ticket, uint32_t num_tiles, int allow_tma, int chain_rows, int options,
                         const uint32_t* __restrict__ d_n = nullptr, const uint2* __restrict__ tile_info = nullptr)
{
    static_assert(THREADS >= k_radix && THREADS % 32 == 0, "one thread per digit");
    static_assert(IPT % 2 == 0, "ranks are packed two per register");
    constexpr bool KEYS_ONLY = (FLAVOR & k_flavor_keys_only) != 0;
    constexpr uint32_t FLIP = (FLAVOR & k_flavor_descending) ? 0xffffffffu : 0u;
    constexpr uint32_t PAD_KEY = ~FLIP; // ranks after every real key of its tile
    using Smem = RingSmem<THREADS, IPT, !KEYS_ONLY>;
    constexpr int WARPS = Smem::WARPS;
    constexpr int TILE = Smem::TILE;
    constexpr int WARP_ELEMS = IPT * 32;
    static_assert(TILE <= 65536, "16-bit tile-local ranks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chain_ctas = chain_rows >= 100 ? 4u : 8u;
    if constexpr (SEG)
    {
        num_tiles = __ldg(d_n);
        n = num_tiles * uint32_t(TILE);
    }
    else if (d_n)
    {
        // *_dyn entry points: the count is device-resident (<= the n the scratch was sized for)
        n = __ldg(d_n);
        num_tiles = (n + uint32_t(TILE) - 1) / uint32_t(TILE);
    }

    for (int i = tid; i < WARPS * k_radix / 4; i += THREADS)
        reinterpret_cast<uint4*>(&s.warp_hist[0][0])[i] = make_uint4(0, 0, 0, 0);

    if (blockIdx.x < chain_ctas)
    {
        __syncthreads();
        uint32_t* totals = reinterpret_cast<uint32_t*>(&s.warp_hist[0][0]);
        switch (chain_rows)
        {
        case 2: chain_cta<WARPS, 2, 1, true>(totals, blockIdx.x, lookback, prefix, num_tiles); break;
        case 4: chain_cta<WARPS, 4, 1, true>(totals, blockIdx.x, lookback, prefix, num_tiles); break;
        case 104: chain_cta<WARPS, 4, 2, true>(totals, blockIdx.x, lookback, prefix, num_tiles); break;
        case 108: chain_cta<WARPS, 8, 2, true>(totals, blockIdx.x, lookback, prefix, num_tiles); break;
        default: chain_cta<WARPS, 8, 1, true>(totals, blockIdx.x, lookback, prefix, num_tiles); break;
        }
        return;
    }

    auto digit_of = [&](uint32_t k) -> uint32_t { return ((k ^ FLIP) >> shift) & mask; };
    // SEG: every tile is a whole bulk copy (padding slots are masked when the keys go to registers)
    auto tma_ok = [&](uint32_t t) -> bool {
        return allow_tma && (SEG || uint64_t(t + 1) * uint32_t(TILE) <= uint64_t(n));
    };

    uint32_t ahead = 0xffffffffu; // thread 0: the ticket drawn for the tile after next
    // options bits 8..: tiles a CTA may draw before it stops drawing and retires (0 = until the tickets run out).  A
    // bounded life lets the block scheduler interleave other streams' kernels (the multi-GPU pipeline) at a granularity
    // of a few tiles; the grid then holds more CTAs than are resident at once, which tickets make safe.
    uint32_t budget = uint32_t(options) >> 8;
    budget = budget ? budget : 0xffffffffu;
    uint64_t policy = 0;
    if (tid == 0)
    {
        mbarrier_init(&s.bar_keys[0], 1);
        mbarrier_init(&s.bar_keys[1], 1);
        mbarrier_init(&s.bar_vals, 1);
        mbarrier_init_fence();
        policy = l2_policy_evict_first();
        const uint32_t t0 = atomicAdd(ticket, 2u); // this CTA's first two tiles
        s.tile_of[0] = t0;
        s.tile_of[1] = t0 + 1;
        if constexpr (SEG)
        {
            s.info_of[0] = t0 < num_tiles ? tile_info[t0] : make_uint2(0, 0);
            s.info_of[1] = t0 + 1 < num_tiles ? tile_info[t0 + 1] : make_uint2(0, 0);
        }
        if (t0 < num_tiles && tma_ok(t0))
        {
            mbarrier_arrive_expect_tx(&s.bar_keys[0], TILE * 4);
            tma_load_1d(s.keys[0], keys_in + size_t(t0) * TILE, TILE * 4, &s.bar_keys[0], policy);
            if constexpr (!KEYS_ONLY)
            {
                mbarrier_arrive_expect_tx(&s.bar_vals, TILE * 4);
                tma_load_1d(s.vals, vals_in + size_t(t0) * TILE, TILE * 4, &s.bar_vals, policy);
            }
        }
        if (t0 + 1 < num_tiles && tma_ok(t0 + 1))
        {
            mbarrier_arrive_expect_tx(&s.bar_keys[1], TILE * 4);
            tma_load_1d(s.keys[1], keys_in + size_t(t0 + 1) * TILE, TILE * 4, &s.bar_keys[1], policy);
            if constexpr (!KEYS_ONLY)
                tma_prefetch_l2_1d(vals_in + size_t(t0 + 1) * TILE, TILE * 4);
        }
        ahead = budget > 2 ? atomicAdd(ticket, 1u) : 0xffffffffu;
        budget = budget > 3 ? budget - 3 : 0;
    }
    __syncthreads();

    uint32_t cur = s.tile_of[0];
    if (cur >= num_tiles)
        return;

    const uint32_t my_off = warp * WARP_ELEMS + lane; // + i * 32   (warp-striped: input order inside the warp)
    uint32_t* wh = s.warp_hist[warp];
    uint32_t key[IPT];
    uint32_t kparity = 0, vparity = 0; // bit b: phase the next wait on bar_keys[b] / bar_vals looks for

    // keys of tile t (ring slot b) -> registers, and into the warp's digit counters
    // number of real pairs of tile t (info: its tile_info word, SEG only)
    auto valid_of = [&](uint32_t t, uint32_t info_x) -> uint32_t {
        if constexpr (SEG)
            return info_x & 0xffffffu;
        else
        {
            const uint32_t base = t * uint32_t(TILE);
            return n - base < uint32_t(TILE) ? n - base : uint32_t(TILE);
        }
    };
    auto load_and_count = [&](uint32_t b, uint32_t t, uint32_t info_x) {
        const uint32_t valid = valid_of(t, info_x);
        if (tma_ok(t))
        {
            mbarrier_wait(&s.bar_keys[b], (kparity >> b) & 1u);
            kparity ^= 1u << b;
#pragma unroll
            for (int i = 0; i < IPT; i++)
                key[i] = s.keys[b][my_off + i * 32];
            if constexpr (SEG)
            {
                if (valid != uint32_t(TILE)) // the last tile of a segment: what lies behind it is not data
                {
#pragma unroll
                    for (int i = 0; i < IPT; i++)
                        key[i] = my_off + i * 32 < valid ? key[i] : PAD_KEY;
                }
            }
        }
        else
        {
            // the last, partial tile and 16-byte-misaligned inputs: straight from global memory; slots past the
            // end hold the largest key — they rank after every real key of the tile and are never written back
            const uint32_t base = t * uint32_t(TILE);
#pragma unroll
            for (int i = 0; i < IPT; i++)
                key[i] = my_off + i * 32 < valid ? keys_in[base + my_off + i * 32] : PAD_KEY;
        }
        // A digit shared by the whole warp would be a 32-way same-address atomic.  Probe the first key: a warp
        // that looks skewed checks every key and counts warp-uniform digits once.
        const uint32_t d_first = digit_of(key[0]);
        if (__all_sync(k_full_mask, d_first == __shfl_sync(k_full_mask, d_first, 0)))
        {
#pragma unroll
            for (int i = 0; i < IPT; i++)
            {
                const uint32_t d = digit_of(key[i]);
                if (__all_sync(k_full_mask, d == __shfl_sync(k_full_mask, d, 0)))
                {
                    if (lane == 0)
                        wh[d] += 32;
                }
                else
                    atomicAdd(&wh[d], 1u);
            }
        }
        else
        {
#pragma unroll
            for (int i = 0; i < IPT; i++)
                atomicAdd(&wh[digit_of(key[i])], 1u);
        }
    };

    // digit threads: the tile's digit counts summed over the warps, PUBLISHED for the chain CTAs at once
    auto publish = [&](uint32_t t, uint32_t info_x) -> uint32_t {
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++)
            tot += s.warp_hist[w][tid];
        const uint32_t valid_t = valid_of(t, info_x);
        // padding slots all carry the digit of the padding key
        const uint32_t count_valid = tot - (tid == digit_of(PAD_KEY) ? uint32_t(TILE) - valid_t : 0u);
        st_relaxed_u32(&lookback[size_t(t) * k_radix + tid], k_lb_local | count_valid);
        return tot;
    };

    uint2 cur_info = make_uint2(0, 0); // SEG: tile_info of the current tile
    if constexpr (SEG)
        cur_info = s.info_of[0];
    load_and_count(0, cur, cur_info.x);
    __syncthreads();
    uint32_t total = 0; // digit threads: digit count of the current tile (padding included)
    if (tid < k_radix)
        total = publish(cur, cur_info.x);

    for (uint32_t it = 0;; it++)
    {
        const uint32_t b = it & 1u;
        const uint32_t tile = cur;
        const uint32_t tile_base = tile * uint32_t(TILE);
        const uint32_t valid = valid_of(tile, cur_info.x);
        const bool full = valid == uint32_t(TILE);
        const bool use_tma = tma_ok(tile);
        uint32_t* skeys = s.keys[b];
        // thread 0: tile_info of the ticket drawn an iteration ago, requested now, stored with the refill (step 6)
        uint2 ahead_info = make_uint2(0, 0);
        if constexpr (SEG)
        {
            if (tid == 0 && ahead < num_tiles)
                ahead_info = tile_info[ahead];
        }

        // ---- 1. per digit: scan of the tile's digit counts; slot offsets of each warp
        uint32_t inc = 0;
        if (tid < k_radix)
        {
            inc = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
                if (lane >= unsigned(o))
                    inc += t;
            }
            if (lane == 31)
                s.scan[warp] = inc;
        }
        __syncthreads();
        if (tid < k_radix)
        {
            uint32_t tile_start = inc - total;
            for (unsigned w = 0; w < warp; w++)
                tile_start += s.scan[w];
            s.tile_start[tid] = tile_start;
            uint32_t running = tile_start;
#pragma unroll
            for (int w = 0; w < WARPS; w++)
            {
                const uint32_t c = s.warp_hist[w][tid];
                s.warp_hist[w][tid] = running;
                running += c;
            }
        }
        __syncthreads();

        // ---- 2. rank + scatter keys (in place: every key of the tile is in registers)
        uint32_t rank2[IPT / 2]; // two 16-bit tile-sorted slots per register
        {
            const uint32_t lt = lanemask_lt();
#pragma unroll
            for (int i = 0; i < IPT; i++)
            {
                const uint32_t d = digit_of(key[i]);
                const uint32_t peers = match_digit<MODE>(d);
                uint32_t before;
                if constexpr (RATOM)
                {
                    const uint32_t lower = peers & lt;
                    before = 0;
                    if (lower == 0)
                        before = atomicAdd(&wh[d], uint32_t(__popc(peers)));
                    before = __shfl_sync(k_full_mask, before, __ffs(peers) - 1);
                }
                else
                {
                    before = wh[d];
                    __syncwarp();
                    wh[d] = before + __popc(peers); // every peer stores the same value
                    __syncwarp();
                }
                const uint32_t r = before + __popc(peers & lt);
                skeys[r] = key[i];
                if (i & 1)
                    rank2[i / 2] |= r << 16;
                else
                    rank2[i / 2] = r;
            }
            // this warp's counters are free again: cleared for the early counts of the next tile
            __syncwarp();
            reinterpret_cast<uint4*>(wh)[lane] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(wh)[lane + 32] = make_uint4(0, 0, 0, 0);
        }

        // ---- 3. early counts of the next tile (its keys stay in registers until iteration it + 1 ranks them).  They are
        //         published BEFORE this tile's look-back: a tile's counts never wait for anything but its own bulk copy,
        //         so the chain CTAs always find them there.
        const uint32_t nxt = s.tile_of[b ^ 1u];
        const bool has_next = nxt < num_tiles;
        uint2 nxt_info = make_uint2(0, 0);
        if constexpr (SEG)
            nxt_info = s.info_of[b ^ 1u];
        __syncwarp(); // the cleared counters
        if (has_next)
            load_and_count(b ^ 1u, nxt, nxt_info.x);

        // ---- 4. values -> registers; next tile's counts published; this digit's count in all earlier tiles (one row,
        //         written by the chain CTA); values -> tile-sorted slot
        {
            uint32_t val[KEYS_ONLY ? 2 : IPT];
            if constexpr (!KEYS_ONLY)
            {
                if (use_tma)
                {
                    mbarrier_wait(&s.bar_vals, vparity);
                    vparity ^= 1u;
#pragma unroll
                    for (int i = 0; i < IPT; i++)
                        val[i] = s.vals[my_off + i * 32];
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < IPT; i++)
                        val[i] = my_off + i * 32 < valid ? vals_in[tile_base + my_off + i * 32] : 0u;
                }
            }
            __syncthreads(); // all values are in registers; the next tile's counts are final; keys are tile-sorted
            if (tid < k_radix)
            {
                if (has_next)
                    total = publish(nxt, nxt_info.x);
                uint32_t exclusive = 0;
                const uint32_t first = SEG ? cur_info.y : 0u; // first tile of the sequence this tile belongs to
                if (tile > first && !(options & k_opt_no_lookback)) // k_opt_no_lookback: timing experiments only
                {
                    const uint32_t* p = prefix + size_t(tile - 1) * k_radix + tid;
                    uint32_t x = ld_relaxed_u32(p);
                    while ((x & k_lb_inclusive) == 0)
                        x = ld_relaxed_u32(p);
                    exclusive = x & ~k_lb_inclusive;
                    if (SEG && first > 0)
                    {
                        // the running prefix does not restart at a segment: take off what precedes the segment
                        const uint32_t* q = prefix + size_t(first - 1) * k_radix + tid;
                        uint32_t y = ld_relaxed_u32(q);
                        while ((y & k_lb_inclusive) == 0)
                            y = ld_relaxed_u32(q);
                        exclusive -= y & ~k_lb_inclusive;
                    }
                }
                const uint32_t seg_row = SEG ? (cur_info.x >> 24) * uint32_t(k_radix) : 0u;
                s.gbase[tid] = digit_offset[seg_row + tid] + exclusive - s.tile_start[tid];
            }
            if constexpr (!KEYS_ONLY)
            {
#pragma unroll
                for (int i = 0; i < IPT; i += 2)
                {
                    s.vals[rank2[i / 2] & 0xffffu] = val[i];
                    s.vals[rank2[i / 2] >> 16] = val[i + 1];
                }
            }
        }
        __syncthreads(); // tile-sorted keys and values, gbase

        // ---- 5. out: consecutive threads write consecutive addresses inside each digit run
        if (full)
        {
#pragma unroll
            for (int k = 0; k < IPT; k++)
            {
                const uint32_t p = tid + k * THREADS;
                const uint32_t kk = skeys[p];
                if constexpr (KEYS_ONLY)
                    keys_out[s.gbase[digit_of(kk)] + p] = kk;
                else
                {
                    const uint32_t vv = s.vals[p];
                    const uint32_t dst = s.gbase[digit_of(kk)] + p;
                    keys_out[dst] = kk;
                    vals_out[dst] = vv;
                }
            }
        }
        else
        {
            for (uint32_t p = tid; p < valid; p += THREADS)
            {
                const uint32_t kk = skeys[p];
                const uint32_t dst = s.gbase[digit_of(kk)] + p;
                keys_out[dst] = kk;
                if constexpr (!KEYS_ONLY)
                    vals_out[dst] = s.vals[p];
            }
        }
        if (!has_next)
            break;
        // this thread's generic-proxy accesses to the staging buffers (tile-sorted scatters, read-back) are ordered
        // before the bulk copies that thread 0 issues into the same buffers after the barrier
        fence_proxy_async_smem();
        __syncthreads(); // the tile has left shared memory

        // ---- 6. refill: values of the next tile, keys of the tile after next
        if (tid == 0)
        {
            fence_proxy_async_smem(); // the generic-proxy reads above precede the bulk copies' writes
            if constexpr (!KEYS_ONLY)
            {
                if (tma_ok(nxt))
                {
                    mbarrier_arrive_expect_tx(&s.bar_vals, TILE * 4);
                    tma_load_1d(s.vals, vals_in + size_t(nxt) * TILE, TILE * 4, &s.bar_vals, policy);
                }
            }
            s.tile_of[b] = ahead;
            if constexpr (SEG)
                s.info_of[b] = ahead_info;
            if (ahead < num_tiles)
            {
                if (tma_ok(ahead))
                {
                    mbarrier_arrive_expect_tx(&s.bar_keys[b], TILE * 4);
                    tma_load_1d(s.keys[b], keys_in + size_t(ahead) * TILE, TILE * 4, &s.bar_keys[b], policy);
                    if constexpr (!KEYS_ONLY)
                        tma_prefetch_l2_1d(vals_in + size_t(ahead) * TILE, TILE * 4);
                }
                ahead = budget ? atomicAdd(ticket, 1u) : 0xffffffffu;
                budget -= budget ? 1u : 0u;
            }
        }
        cur = nxt;
        cur_info = nxt_info;
    }
}
