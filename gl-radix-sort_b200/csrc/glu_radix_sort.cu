// glu_radix_sort.cu — glu_radix_sort_u32kv(): the B200 replacement for glu::RadixSort::operator()
// (glu/RadixSort.hpp:273-334; counting shader :11-58, reordering shader :60-183).
//
// The reference sorts with 8 passes over 4-bit digits; every pass is a counting dispatch (one global
// atomic per key), a 16-partition Blelloch scan of the per-block counts (2*log2(blocks) dispatches) and
// a reordering dispatch that runs 16 sequential 1024-wide shared-memory scans per block: 176..304
// dispatches per sort and a sync-bound ~53 Mpairs/s plateau.  Here the same stable LSD sort is
//   1 memset + 1 histogram kernel + ceil(bits/8) "onesweep" kernels   (68 B of HBM traffic per pair):
//   * histogram_kernel reads the keys ONCE (128-bit streaming loads) and builds the 256-bin
//     histograms of all digit places in shared memory with one private copy of every bin per lane
//     (no bank conflicts, no same-address pile-up: skewed inputs cost the same as uniform ones);
//     the last CTA turns the histograms into exclusive digit offsets;
//   * onesweep_kernel (one launch per 8-bit digit) processes one tile per CTA: keys and values are staged
//     by TMA bulk copies, keys are ranked with warp-wide digit matching against per-warp digit counters
//     (stable: warp-striped order is input order), the per-tile digit counts are published at once and
//     chained across tiles by a few dedicated CHAIN CTAs (decoupled look-back turned inside out: a tile
//     reads ONE prefix row instead of walking back over its predecessors; status + count in one 32-bit
//     word per (tile, digit)), keys and values are reordered through shared memory so that every digit
//     run leaves the SM as one contiguous, coalesced store burst.  Tile ids are blockIdx.x (an atomic
//     ticket is optional, GLU_SORT_OPTIONS bit 1).
// Results are bit-identical to std::stable_sort of the (key, value) pairs by key — and therefore to
// the reference's 8 x 4-bit passes, which are stable as well (oracle/glu_oracle.cpp radix_sort_glsl).
#include <cstdlib>

#include "glu_common.cuh"

namespace glu_b200
{
    namespace
    {
        constexpr int k_radix = 256;
        constexpr int k_max_passes = 4;
        constexpr uint32_t k_lb_local = 1u << 30;     // counts row: tile-local digit count published (bits 0..29)
        constexpr uint32_t k_lb_inclusive = 1u << 31; // prefix row: inclusive count over tiles 0..t (bits 0..30)
        constexpr uint32_t k_lb_value_mask = (1u << 30) - 1u;
        constexpr int k_opt_no_lookback = 1, k_opt_ticket = 2, k_opt_copy4 = 4, k_opt_copy16 = 8; // GLU_SORT_OPTIONS bits
        // onesweep_kernel FLAVOR bits (glu_radix_sort_u32_ex): no value array; digits complemented (descending order)
        constexpr int k_flavor_keys_only = 1, k_flavor_descending = 2;
        // 31-bit running digit counts in the prefix rows (bit 31 is the flag), 32-bit element indices
        constexpr size_t k_max_count = (size_t(1) << 31) - 1;

        struct PassPlan
        {
            int num_passes;
            uint32_t begin_bit; // first key bit that takes part (0 for glu::RadixSort; glu_radix_sort_u32_ex)
            uint32_t key_mask;  // (key >> begin_bit) & key_mask = the bits that take part (glu/RadixSort.hpp:331 num_steps)
            uint32_t shift[k_max_passes];
            uint32_t mask[k_max_passes];
        };

        // key bits [begin_bit, begin_bit + bits) take part, least significant digit first
        PassPlan make_bit_plan(unsigned begin_bit, int bits)
        {
            PassPlan p{};
            p.begin_bit = begin_bit;
            p.key_mask = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
            p.num_passes = (bits + 7) / 8;
            for (int i = 0; i < p.num_passes; i++)
            {
                int b = bits - 8 * i < 8 ? bits - 8 * i : 8;
                p.shift[i] = begin_bit + 8u * i;
                p.mask[i] = (1u << b) - 1u;
            }
            return p;
        }

        PassPlan make_pass_plan(size_t num_steps)
        {
            return make_bit_plan(0u, (num_steps == 0 || num_steps >= 8) ? 32 : int(4 * num_steps));
        }

        // ------------------------------------------------------------------------------------ histogram

        constexpr int k_hist_threads = 1024;
        constexpr int k_hist_unroll = 4;
        constexpr int k_hist_blocks_per_sm = 1;
        constexpr int k_hist_copies = 32; // one private copy of every bin per lane

        // hist: [k_max_passes][256] zero-initialised; on exit hist holds EXCLUSIVE digit offsets
        // (make_offsets) or the plain counts.  Digit place p of a key is ((key >> pre_shift) & key_mask) >> 8p.
        //
        // Shared-memory atomics are the cost here (one per key and digit place), and what makes them slow is
        // lanes of a warp meeting in a bank.  So every bin has 32 copies, one per LANE, in 32 consecutive
        // words: lane l only ever touches bank l — no bank conflicts and no same-address pile-up inside a warp
        // whatever the key distribution (uniform, constant and skewed digit places cost the same), at 32 KB
        // of shared memory per digit place (one 1024-thread CTA per SM).
        __global__ void __launch_bounds__(k_hist_threads, 1)
            histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ d_n,
                             uint32_t head, int num_passes, uint32_t pre_shift, uint32_t key_mask, uint32_t* hist,
                             uint32_t* ticket, int make_offsets, int descending)
        {
            // d_n: the count lives in device memory (written by an earlier kernel of the stream, *_dyn entry points)
            if (d_n)
                n = __ldg(d_n);
            head = head < n ? head : n;                // keys before the first 16-byte boundary
            const uint32_t n_units = (n - head) / 4;   // 128-bit units of the aligned body
            extern __shared__ __align__(16) uint32_t s_hist[]; // [num_passes][k_radix][k_hist_copies]
            __shared__ uint32_t s_scan[k_hist_threads / 32];
            __shared__ bool s_is_last;

            for (int i = threadIdx.x; i < num_passes * k_radix * k_hist_copies / 4; i += k_hist_threads)
                reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();

            const unsigned lane = threadIdx.x & 31;
            uint32_t* mine = s_hist + lane;
            const uint4* body = reinterpret_cast<const uint4*>(keys + head);
            const uint32_t stride = gridDim.x * k_hist_threads;
            for (uint32_t v0 = blockIdx.x * k_hist_threads + threadIdx.x; v0 < n_units; v0 += stride * k_hist_unroll)
            {
                uint4 k[k_hist_unroll];
                bool ok[k_hist_unroll];
#pragma unroll
                for (int u = 0; u < k_hist_unroll; u++)
                {
                    const uint64_t v = uint64_t(v0) + uint64_t(u) * stride;
                    ok[u] = v < n_units;
                    k[u] = ok[u] ? ld_stream_v4(body + v) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < k_hist_unroll; u++)
                {
                    if (!ok[u])
                        continue;
                    const uint32_t kk[4] = {(k[u].x >> pre_shift) & key_mask, (k[u].y >> pre_shift) & key_mask,
                                            (k[u].z >> pre_shift) & key_mask, (k[u].w >> pre_shift) & key_mask};
#pragma unroll
                    for (int p = 0; p < k_max_passes; p++)
                    {
                        if (p >= num_passes) // uniform
                            break;
#pragma unroll
                        for (int c = 0; c < 4; c++)
                        {
                            // byte p of the key (PRMT), scaled address (LEA), atomic: three instructions per count
                            const uint32_t d = __byte_perm(kk[c], 0u, 0x4440u + p);
                            atomicAdd(mine + p * k_radix * k_hist_copies + d * k_hist_copies, 1u);
                        }
                    }
                }
            }
            // unaligned head and sub-vector tail (at most 3 keys each)
            if (blockIdx.x == 0 && threadIdx.x < 32)
            {
                const uint32_t tail_begin = head + n_units * 4;
                const uint32_t idx = lane < head ? lane : tail_begin + (lane - head);
                const bool ok = idx < n && lane < head + 3;
                const uint32_t key = ok ? ((keys[idx] >> pre_shift) & key_mask) : 0;
                if (ok)
                    for (int p = 0; p < num_passes; p++)
                        atomicAdd(mine + ((p * k_radix + ((key >> (8 * p)) & 0xffu)) * k_hist_copies), 1u);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < num_passes * k_radix; i += k_hist_threads)
            {
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < k_hist_copies; j++)
                    c += s_hist[i * k_hist_copies + ((j + lane) & (k_hist_copies - 1))]; // rotated: bank-conflict free
                if (c)
                    atomicAdd(&hist[i], c);
            }
            if (!make_offsets) // glu_radix_histogram_u32: raw counts
                return;

            // last CTA: counts -> exclusive offsets, one digit place at a time.  descending: the passes partition by
            // the COMPLEMENTED digit (mask ^ digit), so entry t is the number of keys whose complemented digit is < t
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0)
                s_is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
            __syncthreads();
            if (!s_is_last)
                return;
            __threadfence();
            const unsigned warp = threadIdx.x >> 5;
            for (int p = 0; p < num_passes; p++)
            {
                uint32_t c = 0, inc = 0;
                if (threadIdx.x < k_radix)
                {
                    const uint32_t flip = descending ? ((key_mask >> (8 * p)) & 0xffu) : 0u;
                    c = __ldcg(&hist[p * k_radix + (threadIdx.x ^ flip)]); // all reads precede the barrier, all writes follow it
                    inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
                        if (lane >= unsigned(o))
                            inc += t;
                    }
                    if (lane == 31)
                        s_scan[warp] = inc;
                }
                __syncthreads();
                if (threadIdx.x < k_radix)
                {
                    uint32_t off = 0;
                    for (unsigned w = 0; w < warp; w++)
                        off += s_scan[w];
                    hist[p * k_radix + threadIdx.x] = off + inc - c;
                }
                __syncthreads();
            }
        }

        // ------------------------------------------------------------------------------------ onesweep

        enum RankMode
        {
            Rank_Match = 0, // match.any.sync
            Rank_Ballot = 1 // one ballot per digit bit
        };

        // One step of the ballot match: lanes whose digit agrees with mine in bit BIT stay in `peers`.
        // Hand-scheduled as ~3 SASS instructions per bit (R2P for 7 bits at once, VOTE, SEL, LOP3);
        // the C++ formulation `peers &= bit ? vote : ~vote` compiles to 6-7.
#define GLU_MATCH_BIT(BIT)                                                                                             \
    "and.b32 t, %1, " #BIT ";\n"                                                                                       \
    "setp.ne.u32 p, t, 0;\n"                                                                                           \
    "vote.sync.ballot.b32 v, p, 0xffffffff;\n"                                                                         \
    "selp.b32 m, 0xffffffff, 0, p;\n"                                                                                  \
    "lop3.b32 %0, %0, v, m, 0x90;\n" /* peers & ~(v ^ m) */

        template<int MODE> __device__ __forceinline__ uint32_t match_digit(uint32_t d)
        {
            if (MODE == Rank_Match)
                return __match_any_sync(k_full_mask, d);
            uint32_t peers;
            asm volatile("{\n"
                         ".reg .pred p;\n"
                         ".reg .b32 v, t, m;\n"
                         "mov.b32 %0, 0xffffffff;\n" GLU_MATCH_BIT(1) GLU_MATCH_BIT(2) GLU_MATCH_BIT(4) GLU_MATCH_BIT(8)
                             GLU_MATCH_BIT(16) GLU_MATCH_BIT(32) GLU_MATCH_BIT(64) GLU_MATCH_BIT(128) "}\n"
                         : "=&r"(peers)
                         : "r"(d));
            return peers;
        }

        // same for digits < 16 (the destination ids of glu_radix_partition_by_dest_u32kv): four rounds
        __device__ __forceinline__ uint32_t match_low4(uint32_t d)
        {
            uint32_t peers;
            asm volatile("{\n"
                         ".reg .pred p;\n"
                         ".reg .b32 v, t, m;\n"
                         "mov.b32 %0, 0xffffffff;\n" GLU_MATCH_BIT(1) GLU_MATCH_BIT(2) GLU_MATCH_BIT(4) GLU_MATCH_BIT(8) "}\n"
                         : "=&r"(peers)
                         : "r"(d));
            return peers;
        }
#undef GLU_MATCH_BIT

        
        // PEER = true: every digit run goes to its own destination pointer (possibly in another GPU's memory,
        // glu_radix_partition_u32kv) instead of one output array.
        template<bool PEER> struct SweepDst
        {
        };
        template<> struct SweepDst<true>
        {
            uint32_t* key[k_radix]; // address of tile-sorted slot 0, per digit
            uint32_t* val[k_radix];
            uint8_t lut[k_radix];   // DEST: key digit -> destination id (the "digit" the pass partitions by)
        };

        template<int RANK_THREADS, int IPT, bool PEER = false, bool HAS_VALS = true> struct SweepSmem
        {
            static constexpr int WARPS = RANK_THREADS / 32; // ranking warps
            static constexpr int TILE = RANK_THREADS * IPT;
            alignas(128) uint32_t keys[TILE];   // TMA destination (input order), then tile-sorted keys
            alignas(128) uint32_t vals[HAS_VALS ? TILE : 32]; // TMA destination (input order), then tile-sorted values
            alignas(16) uint32_t warp_hist[WARPS][k_radix]; // per-warp digit counts, then running slot offsets
            uint32_t gbase[k_radix];            // global index of tile-sorted slot 0, per digit
            uint32_t tile_start[k_radix];       // first tile-sorted slot of each digit
            uint32_t scan[8];
            SweepDst<PEER> dst;
            alignas(8) uint64_t bar_keys;       // mbarriers completed by the bulk copies
            alignas(8) uint64_t bar_vals;
            uint32_t tile;
        };

        // The CHAIN CTAs of a onesweep pass (see onesweep_kernel).  Chain CTA c turns the tiles' digit counts
        // ("count rows") of digits [32 * GROUPS * c, ...) into running prefixes ("prefix rows").  A group of WPG
        // warps owns 32 digits (one per lane) and walks the rows in batches of WPG * ROWS: warp w takes ROWS
        // consecutive rows of the batch, requests them all at once (late rows are asked for again TOGETHER: one L2
        // round trip per polling round however many are late), takes the running total of everything before its
        // rows from its predecessor warp through shared memory (a ring: warp 0 continues from the last warp of the
        // previous batch) and hands its own running total on.  ORDERED (the ring kernel, whose CTAs hold tickets of
        // tiles they have not counted yet): prefix row t is written the moment count rows <= t are known, so it NEVER
        // depends on a later tile; otherwise a warp simply waits for all rows of its slice.  The next batch's requests
        // are in flight meanwhile.
        template<int WARPS, int ROWS, int GROUPS, bool ORDERED = false>
        __device__ __noinline__ void chain_cta(uint32_t* smem, uint32_t chain_id, const uint32_t* lookback,
                                               uint32_t* prefix, uint32_t num_tiles)
        {
            constexpr int WPG = WARPS / GROUPS; // warps per 32-digit group
            constexpr int BATCH = WPG * ROWS;   // rows per batch
            static_assert(WPG >= 2, "the carry ring needs two warps");
            // smem was zeroed by the caller: carry values, then one "batches handed on" counter per warp
            uint32_t(*carry)[GROUPS][WPG][32] = reinterpret_cast<uint32_t(*)[GROUPS][WPG][32]>(smem);
            volatile uint32_t* seq = smem + 2 * GROUPS * WPG * 32;
            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (warp >= GROUPS * WPG)
                return;
            const unsigned g = warp / WPG, w = warp % WPG;
            const uint32_t d = (chain_id * GROUPS + g) * 32 + lane;
            const uint32_t* col = lookback + d;
            uint32_t p[ROWS], q[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; j++)
                q[j] = w * ROWS + j < num_tiles ? ld_relaxed_u32(col + size_t(w * ROWS + j) * k_radix) : k_lb_local;
            if constexpr (!ORDERED)
            {
                // one tile per CTA (onesweep_kernel): every tile publishes its counts before it waits for anything, so a
                // warp may wait for ALL rows of its slice first — the shortest hand-off path
                uint32_t batch = 0;
                for (uint32_t t0 = 0; t0 < num_tiles; t0 += BATCH, batch++)
                {
                    const uint32_t r0 = t0 + w * ROWS;
#pragma unroll
                    for (int j = 0; j < ROWS; j++)
                        p[j] = q[j];
                    while (true)
                    {
                        uint32_t all = k_lb_local;
#pragma unroll
                        for (int j = 0; j < ROWS; j++)
                            all &= p[j];
                        if (all & k_lb_local)
                            break;
#pragma unroll
                        for (int j = 0; j < ROWS; j++)
                            if ((p[j] & k_lb_local) == 0)
                                p[j] = ld_relaxed_u32(col + size_t(r0 + j) * k_radix);
                    }
#pragma unroll
                    for (int j = 0; j < ROWS; j++)
                        q[j] = r0 + BATCH + j < num_tiles ? ld_relaxed_u32(col + size_t(r0 + BATCH + j) * k_radix) : k_lb_local;
                    uint32_t run = 0;
#pragma unroll
                    for (int j = 0; j < ROWS; j++)
                    {
                        run += p[j] & k_lb_value_mask;
                        p[j] = run;
                    }
                    // running total of all rows before mine
                    uint32_t in = 0;
                    if (w > 0 || batch > 0)
                    {
                        const unsigned pw = w > 0 ? w - 1 : WPG - 1;
                        const uint32_t pb = w > 0 ? batch : batch - 1;
                        while (seq[g * WPG + pw] < pb + 1)
                        {
                        }
                        __threadfence_block();
                        in = *const_cast<volatile uint32_t*>(&carry[pb & 1][g][pw][lane]);
                    }
                    *const_cast<volatile uint32_t*>(&carry[batch & 1][g][w][lane]) = in + run;
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0)
                        seq[g * WPG + w] = batch + 1;
#pragma unroll
                    for (int j = 0; j < ROWS; j++)
                        if (r0 + j < num_tiles)
                            st_relaxed_u32(prefix + size_t(r0 + j) * k_radix + d, k_lb_inclusive | (in + p[j]));
                }
                return;
            }
            uint32_t batch = 0;
            for (uint32_t t0 = 0; t0 < num_tiles; t0 += BATCH, batch++)
            {
                const uint32_t r0 = t0 + w * ROWS;
#pragma unroll
                for (int j = 0; j < ROWS; j++)
                    p[j] = q[j];
                // the next batch's rows are requested before this one is combined
#pragma unroll
                for (int j = 0; j < ROWS; j++)
                    q[j] = r0 + BATCH + j < num_tiles ? ld_relaxed_u32(col + size_t(r0 + BATCH + j) * k_radix) : k_lb_local;
                // Wait for BOTH the predecessor's running total (`in`, through shared memory) and my rows (global
                // memory; late rows are asked for again TOGETHER: one L2 round trip per round however many are late) —
                // at the same time, not one after the other.  While waiting, a row whose predecessors are all known is
                // published at once: prefix row t depends on count rows <= t ONLY (a tile of the ring kernel waits for
                // prefix row t - 1 before it publishes the counts of its next tile, which may be one of my later rows).
                uint32_t in = 0, run = 0;
                unsigned published = 0; // rows [0, published) of my slice are written
                bool have_in = (w == 0 && batch == 0);
                const unsigned pw = w > 0 ? w - 1 : WPG - 1;
                const uint32_t pb = w > 0 ? batch : batch - 1;
                while (true)
                {
                    if (!have_in && seq[g * WPG + pw] >= pb + 1) // warp-uniform
                    {
                        __threadfence_block();
                        in = *const_cast<volatile uint32_t*>(&carry[pb & 1][g][pw][lane]);
                        have_in = true;
                    }
                    uint32_t all = k_lb_local;
#pragma unroll
                    for (int j = 0; j < ROWS; j++)
                        all &= p[j];
                    const bool rows_ready = (all & k_lb_local) != 0;
                    if (have_in && __all_sync(k_full_mask, rows_ready))
                        break;
                    if (have_in)
                    {
#pragma unroll
                        for (int j = 0; j < ROWS; j++)
                            if (unsigned(j) == published && (p[j] & k_lb_local) != 0)
                            {
                                run += p[j] & k_lb_value_mask;
                                if (r0 + j < num_tiles)
                                    st_relaxed_u32(prefix + size_t(r0 + j) * k_radix + d, k_lb_inclusive | (in + run));
                                published = unsigned(j) + 1;
                            }
                    }
                    if (!rows_ready)
                    {
#pragma unroll
                        for (int j = 0; j < ROWS; j++)
                            if ((p[j] & k_lb_local) == 0)
                                p[j] = ld_relaxed_u32(col + size_t(r0 + j) * k_radix);
                    }
                }
                // everything is known: the short critical section — hand the running total on, THEN write what is left
                uint32_t sum[ROWS];
                uint32_t total = run;
#pragma unroll
                for (int j = 0; j < ROWS; j++)
                {
                    if (unsigned(j) >= published)
                        total += p[j] & k_lb_value_mask;
                    sum[j] = total;
                }
                *const_cast<volatile uint32_t*>(&carry[batch & 1][g][w][lane]) = in + total;
                __threadfence_block();
                __syncwarp();
                if (lane == 0)
                    seq[g * WPG + w] = batch + 1;
#pragma unroll
                for (int j = 0; j < ROWS; j++)
                    if (unsigned(j) >= published && r0 + j < num_tiles)
                        st_relaxed_u32(prefix + size_t(r0 + j) * k_radix + d, k_lb_inclusive | (in + sum[j]));
            }
        }

        // One tile per CTA.  digit(key) = (key >> shift) & mask; keys with equal digits keep their order.
        // The first 8 (or 4) CTAs of the grid are not tiles but the pass's CHAIN CTAs (chain_cta above).
        //
        //   1. tile id (blockIdx.x, or a ticket); one thread issues two cp.async.bulk (TMA) copies: the tile's keys and
        //      values land in shared memory asynchronously, completion on an mbarrier each (the last,
        //      partial tile and 16-byte-misaligned inputs take a cooperative ld/st path instead);
        //   2. EARLY COUNTS: every ranking thread takes IPT keys warp-striped (slot = warp*IPT*32 + i*32
        //      + lane, i.e. input order inside the warp) and bumps its warp's private digit counters;
        //      one thread per digit then sums the warps, PUBLISHES the tile's digit count for look-back
        //      right away (successor tiles need it long before this tile has ranked anything), scans the
        //      256 counts and turns the counters into per-(warp, digit) running slot offsets;
        //   3. ranking: peers = lanes of the warp holding the same digit (one ballot per digit bit);
        //      slot = offset[warp][digit] + popc(peers below me); the key goes straight to its
        //      tile-sorted slot (in place: every key is in registers by now);
        //   4. MEANWHILE the chain CTAs turn the published counts into prefix rows; thread d of the tile
        //      reads prefix[tile - 1][d] (one word, written by the chain) and leaves gbase[digit] = global
        //      index of tile-sorted slot 0;
        //   5. values: staging buffer -> registers -> tile-sorted slot (in place); then slot p of both
        //      arrays goes to global[gbase[digit(key_p)] + p] — neighbouring threads, neighbouring
        //      addresses inside every digit run.
        //
        // PEER: digit run d goes to key_dst[d] / val_dst[d] instead of one output array.  DEST (with PEER): the pass
        // partitions by dest_lut[digit] (< 16 destinations) instead of by the digit itself, so a tile leaves as a
        // handful of long runs — what remote (NVLink) stores want.
        //
        // FLAVOR (glu_radix_sort_u32_ex): k_flavor_keys_only — there is no value array (vals_in / vals_out are not
        // touched); k_flavor_descending — the pass partitions by the complemented digit, which sorts descending and
        // keeps equal keys in input order.  FLAVOR 0 compiles to exactly the code it was before the flavours existed.
        //
        // SEG (glu_radix_sort_seg.cuh): many independent segments in one launch — tile t is elements [t * TILE, (t + 1) *
        // TILE) of a tile-aligned layout, tile_info[t] = {valid | segment << 24, first tile of the segment}, digit_offset
        // is [segment][256] incl. the output base, *d_n the number of tiles; see onesweep_ring_kernel.
        template<int RANK_THREADS, int IPT, int MIN_BLOCKS, int MODE, bool PEER = false, bool DEST = false, int FLAVOR = 0,
                 bool SEG = false>
        __global__ void __launch_bounds__(RANK_THREADS, MIN_BLOCKS)
            onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                            uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                            uint32_t shift, uint32_t mask, const uint32_t* __restrict__ digit_offset,
                            uint32_t* lookback, uint32_t* prefix, uint32_t* ticket, uint32_t num_tiles, int allow_tma,
                            int chain_rows, int options, const uint32_t* __restrict__ d_n = nullptr,
                            uint32_t* const* key_dst = nullptr, uint32_t* const* val_dst = nullptr,
                            const uint8_t* __restrict__ dest_lut = nullptr, const uint2* __restrict__ tile_info = nullptr,
                            const uint32_t* __restrict__ tile_map = nullptr)
        {
            static_assert(!DEST || PEER, "DEST is a flavour of PEER");
            static_assert(!SEG || (!PEER && FLAVOR == 0), "SEG is a flavour of the plain key/value pass");
            static_assert(RANK_THREADS >= k_radix && RANK_THREADS % 32 == 0, "one ranking thread per digit");
            static_assert(IPT % 2 == 0, "ranks are packed two per register");
            static_assert(FLAVOR == 0 || !PEER, "the flavours belong to the single-GPU sort");
            constexpr bool KEYS_ONLY = (FLAVOR & k_flavor_keys_only) != 0;
            constexpr uint32_t FLIP = (FLAVOR & k_flavor_descending) ? 0xffffffffu : 0u;
            constexpr uint32_t PAD_KEY = ~FLIP; // ranks after every real key of its tile
            using Smem = SweepSmem<RANK_THREADS, IPT, PEER, !KEYS_ONLY>;
            constexpr int THREADS = RANK_THREADS;
            constexpr int WARPS = Smem::WARPS;
            constexpr int TILE = Smem::TILE;
            constexpr int WARP_ELEMS = IPT * 32;
            static_assert(TILE <= 65536, "16-bit tile-local ranks");
            extern __shared__ __align__(128) unsigned char smem_raw[];
            Smem& s = *reinterpret_cast<Smem*>(smem_raw);

            const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const uint32_t chain_ctas = chain_rows >= 100 ? 4u : 8u; // 64 or 32 digits per chain CTA
            if constexpr (SEG)
            {
                num_tiles = __ldg(d_n);
                n = num_tiles * uint32_t(TILE);
            }
            else if (d_n)
            {
                // *_dyn entry points: the count is device-resident (<= the n the grid and the scratch were sized
                // for); CTAs past the last tile leave at once
                n = __ldg(d_n);
                num_tiles = (n + uint32_t(TILE) - 1) / uint32_t(TILE);
            }
            if (tid == 0)
            {
                mbarrier_init(&s.bar_keys, 1);
                mbarrier_init(&s.bar_vals, 1);
                mbarrier_init_fence();
                // Tile ids: blockIdx.x (default; CTAs are dispatched in blockIdx order, so every CTA a tile waits
                // for is resident or done — the assumption CUB's DeviceScan makes too; saves the L2 round trip
                // in front of the bulk copies) or an atomic ticket (k_opt_ticket: order by construction).
                const uint32_t t = (options & k_opt_ticket) ? atomicAdd(ticket, 1u) : blockIdx.x;
                s.tile = t;
                // a full tile's bulk copies leave the moment the ticket is known (the rest of the CTA is still
                // clearing its counters)
                if (allow_tma && t >= chain_ctas && uint64_t(t - chain_ctas + 1) * uint32_t(TILE) <= uint64_t(n))
                {
                    // SEG with a tile map (glu_radix_sort_u32kv_segmented_runs, first pass): tile t is READ from tile
                    // tile_map[t] of the input arrays
                    uint32_t tb = (t - chain_ctas) * uint32_t(TILE);
                    if constexpr (SEG)
                    {
                        if (tile_map)
                            tb = __ldg(tile_map + (t - chain_ctas)) * uint32_t(TILE);
                    }
                    const uint64_t policy = l2_policy_evict_first();
                    mbarrier_arrive_expect_tx(&s.bar_keys, TILE * 4);
                    tma_load_1d(s.keys, keys_in + tb, TILE * 4, &s.bar_keys, policy);
                    if constexpr (!KEYS_ONLY)
                    {
                        mbarrier_arrive_expect_tx(&s.bar_vals, TILE * 4);
                        tma_load_1d(s.vals, vals_in + tb, TILE * 4, &s.bar_vals, policy);
                    }
                    // L2 prefetch of the tile that will occupy this CTA slot `options >> 8` tiles from now: its bulk
                    // copies then start from L2 instead of paying the loaded-DRAM latency at CTA start
                    uint64_t ahead = uint64_t(t - chain_ctas) + uint32_t(options >> 8);
                    if ((options >> 8) != 0 && (ahead + 1) * uint64_t(TILE) <= uint64_t(n))
                    {
                        if constexpr (SEG)
                        {
                            if (tile_map)
                                ahead = __ldg(tile_map + ahead);
                        }
                        tma_prefetch_l2_1d(keys_in + ahead * TILE, TILE * 4);
                        if constexpr (!KEYS_ONLY)
                            tma_prefetch_l2_1d(vals_in + ahead * TILE, TILE * 4);
                    }
                }
            }
            for (int i = tid; i < WARPS * k_radix / 4; i += THREADS)
                reinterpret_cast<uint4*>(&s.warp_hist[0][0])[i] = make_uint4(0, 0, 0, 0);
            if constexpr (DEST)
            {
                if (tid < k_radix)
                    s.dst.lut[tid] = dest_lut[tid];
            }
            __syncthreads();
            // what the pass partitions by
            auto digit_of = [&](uint32_t k) -> uint32_t {
                if constexpr (DEST)
                    return s.dst.lut[(k >> shift) & mask];
                else
                    return ((k ^ FLIP) >> shift) & mask;
            };
            if (s.tile < chain_ctas)
            {
                // ---- the first tickets = the CHAIN CTAs (the first CTAs to run, hence resident before any
                // tile exists): see chain_cta().  Tiles never walk back over their predecessors: tile t
                // reads ONE row, prefix[t - 1], which trails the publication of count row t - 1 by about
                // one batch.  chain_rows selects the batch shape (rows per lane; >= 100: 64 digits per CTA).
                uint32_t* totals = reinterpret_cast<uint32_t*>(&s.warp_hist[0][0]);
                switch (chain_rows)
                {
                case 2: chain_cta<WARPS, 2, 1>(totals, s.tile, lookback, prefix, num_tiles); break;
                case 4: chain_cta<WARPS, 4, 1>(totals, s.tile, lookback, prefix, num_tiles); break;
                case 8: chain_cta<WARPS, 8, 1>(totals, s.tile, lookback, prefix, num_tiles); break;
                case 104: chain_cta<WARPS, 4, 2>(totals, s.tile, lookback, prefix, num_tiles); break;
                default: chain_cta<WARPS, 8, 2>(totals, s.tile, lookback, prefix, num_tiles); break;
                }
                return;
            }
            const uint32_t tile = s.tile - chain_ctas;
            if (tile >= num_tiles)
                return;
            uint32_t tile_base = tile * uint32_t(TILE); // where the tile is read from
            if constexpr (SEG)
            {
                if (tile_map)
                    tile_base = __ldg(tile_map + tile) * uint32_t(TILE);
            }
            uint2 info = make_uint2(0, 0);
            if constexpr (SEG)
                info = tile_info[tile];
            const uint32_t valid =
                SEG ? (info.x & 0xffffffu) : (n - tile_base < uint32_t(TILE) ? n - tile_base : uint32_t(TILE));
            const bool full = valid == uint32_t(TILE);
            const bool use_tma = (SEG || full) && allow_tma; // SEG: every tile is a whole bulk copy, padding is masked below
            const uint32_t my_off = warp * WARP_ELEMS + lane; // + i * 32   (warp-striped)

            // ---- stage the tile
            if (!use_tma)
            {
                // Slots past the end of the input hold the largest key: they rank after every real key
                // of the tile and are never written back.
                for (uint32_t idx = tid; idx < uint32_t(TILE); idx += THREADS)
                {
                    s.keys[idx] = idx < valid ? keys_in[tile_base + idx] : PAD_KEY;
                    if constexpr (!KEYS_ONLY)
                        s.vals[idx] = idx < valid ? vals_in[tile_base + idx] : 0u;
                }
                __syncthreads();
            }

            // what item i of this thread is partitioned by.  DEST: the padding slots of the partial last tile get the
            // largest destination id whatever dest_lut holds, so they rank after every real pair of the tile (a
            // non-monotone table must not move them into the middle of the tile-sorted order)
            constexpr uint32_t PAD_DEST = 15u;
            auto digit_at = [&](int i, uint32_t k) -> uint32_t {
                if constexpr (DEST)
                    return (full || my_off + uint32_t(i) * 32u < valid) ? digit_of(k) : PAD_DEST;
                else
                    return digit_of(k);
            };
            const uint32_t pad_digit = DEST ? PAD_DEST : digit_of(PAD_KEY);

            // GLU_SORT_OPTIONS bits 2 / 3 (timing experiments only, wrong results): the tile is written straight back —
            // what a pass costs when nothing but its memory traffic is left (bit 2: 4-byte stores like the real
            // write-out, bit 3: 16-byte stores); the look-back protocol still runs so that the chain CTAs terminate
            if constexpr (!PEER && FLAVOR == 0 && !SEG)
            {
                if (options & (k_opt_copy4 | k_opt_copy16))
                {
                    if (use_tma)
                    {
                        mbarrier_wait(&s.bar_keys, 0);
                        mbarrier_wait(&s.bar_vals, 0);
                    }
                    if (tid < k_radix)
                        st_relaxed_u32(&lookback[size_t(tile) * k_radix + tid], k_lb_local | (tid == 0 ? valid : 0u));
                    if ((options & k_opt_copy16) && full)
                    {
                        for (uint32_t p = tid; p < uint32_t(TILE / 4); p += THREADS)
                        {
                            reinterpret_cast<uint4*>(keys_out + tile_base)[p] = reinterpret_cast<const uint4*>(s.keys)[p];
                            reinterpret_cast<uint4*>(vals_out + tile_base)[p] = reinterpret_cast<const uint4*>(s.vals)[p];
                        }
                    }
                    else
                    {
                        for (uint32_t p = tid; p < valid; p += THREADS)
                        {
                            keys_out[tile_base + p] = s.keys[p];
                            vals_out[tile_base + p] = s.vals[p];
                        }
                    }
                    return;
                }
            }

            // ---- early counts: the warp's digit histogram
            uint32_t key[IPT];
            uint32_t* wh = s.warp_hist[warp];
            {
                if (use_tma)
                    mbarrier_wait(&s.bar_keys, 0);
#pragma unroll
                for (int i = 0; i < IPT; i++)
                    key[i] = s.keys[my_off + i * 32];
                if constexpr (SEG)
                {
                    if (!full) // the last tile of a segment: what lies behind it is not data
                    {
#pragma unroll
                        for (int i = 0; i < IPT; i++)
                            key[i] = my_off + i * 32 < valid ? key[i] : PAD_KEY;
                    }
                }
                if constexpr (DEST)
                {
                    // a handful of destinations: plain atomics would pile up on the same few words, so the lanes
                    // holding the same destination are matched and their leader adds the group's size
                    const uint32_t lt_ = lanemask_lt();
#pragma unroll
                    for (int i = 0; i < IPT; i++)
                    {
                        const uint32_t d = digit_at(i, key[i]);
                        const uint32_t peers = match_low4(d);
                        if ((peers & lt_) == 0)
                            atomicAdd(&wh[d], uint32_t(__popc(peers)));
                    }
                }
                else
                {
                    // A digit shared by the whole warp would be a 32-way same-address atomic.  Probe the first
                    // key: a warp that looks skewed checks every key and counts warp-uniform digits once.
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
                }
            }
            __syncthreads(); // every key is in registers; the counts are final

            // ---- per-digit: tile count -> look-back publication; slot offsets of each warp
            uint32_t total = 0, inc = 0;
            if (tid < k_radix)
            {
#pragma unroll
                for (int w = 0; w < WARPS; w++)
                    total += s.warp_hist[w][tid];
                // padding slots all carry the digit of the padding key
                const uint32_t count_valid = total - (tid == pad_digit ? uint32_t(TILE) - valid : 0u);
                st_relaxed_u32(&lookback[size_t(tile) * k_radix + tid], k_lb_local | count_valid);
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

            uint32_t rank2[IPT / 2]; // two 16-bit tile-sorted slots per register
            {
                // ---- rank + scatter keys (in place)
                const uint32_t lt = lanemask_lt();
#pragma unroll
                for (int i = 0; i < IPT; i++)
                {
                    const uint32_t d = digit_at(i, key[i]);
                    const uint32_t peers = DEST ? match_low4(d) : match_digit<MODE>(d);
                    const uint32_t before = wh[d];
                    __syncwarp();
                    wh[d] = before + __popc(peers); // every peer stores the same value
                    __syncwarp();
                    const uint32_t r = before + __popc(peers & lt);
                    s.keys[r] = key[i];
                    if (i & 1)
                        rank2[i / 2] |= r << 16;
                    else
                        rank2[i / 2] = r;
                }
                // ---- values: staging buffer -> registers (the key registers are dead now)
                uint32_t val[KEYS_ONLY ? 2 : IPT];
                if constexpr (!KEYS_ONLY)
                {
                    if (use_tma)
                        mbarrier_wait(&s.bar_vals, 0);
#pragma unroll
                    for (int i = 0; i < IPT; i++)
                        val[i] = s.vals[my_off + i * 32];
                }

                // ---- this digit's count in all earlier tiles: one row, written by the chain CTA
                auto look_back = [&]() {
                if (tid < k_radix)
                {
                    uint32_t exclusive = 0;
                    const uint32_t first = SEG ? info.y : 0u; // first tile of the sequence this tile belongs to
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
                    if constexpr (PEER)
                    {
                        // only table entries that can be a destination are read: 1 << bits pointers (by digit) or
                        // the first 16 (by destination id) — include/glu_b200.h
                        const ptrdiff_t off = ptrdiff_t(exclusive) - ptrdiff_t(s.tile_start[tid]);
                        const bool in_table = DEST ? tid <= PAD_DEST : tid <= mask;
                        s.dst.key[tid] = in_table ? key_dst[tid] + off : nullptr;
                        s.dst.val[tid] = in_table ? val_dst[tid] + off : nullptr;
                    }
                    else
                        s.gbase[tid] = digit_offset[(SEG ? (info.x >> 24) * uint32_t(k_radix) : 0u) + tid] + exclusive -
                                       s.tile_start[tid];
                }
                };
                // The plain key/value pass asks for the prefix row AFTER the values have been scattered: the row trails the
                // publication of the predecessor's counts by a chain batch, so the later a tile asks, the less it waits
                // (4.48 against 4.53 ms per 2^28-pair sort; the SEG pass, which reads two rows, was 2 % slower that way
                // and keeps the early order, like the PEER pass).
                constexpr bool LATE = !KEYS_ONLY && !SEG && !PEER;
                if constexpr (!LATE)
                    look_back();
                if constexpr (!KEYS_ONLY)
                {
                    __syncthreads(); // all values are in registers
#pragma unroll
                    for (int i = 0; i < IPT; i += 2)
                    {
                        s.vals[rank2[i / 2] & 0xffffu] = val[i];
                        s.vals[rank2[i / 2] >> 16] = val[i + 1];
                    }
                }
                if constexpr (LATE)
                    look_back();
            }
            __syncthreads(); // tile-sorted keys and values, gbase

            // ---- out: consecutive threads write consecutive addresses inside each digit run
            if constexpr (DEST)
            {
                // A tile leaves as <= 16 long runs, most of them to peer GPUs over NVLink.  Every warp-wide store is
                // made to cover ONE 128-byte line of its destination (the window of 32 slots is shifted by the run's
                // misalignment; only the first and last line of a run are partial): a line that straddles two warps'
                // windows would cross the link as two byte-masked packets instead of one full one.
                for (uint32_t d = 0; d <= PAD_DEST; d++)
                {
                    const uint32_t begin = s.tile_start[d];
                    uint32_t end = s.tile_start[d + 1];
                    end = end < valid ? end : valid; // padding slots rank last and are never written
                    if (begin >= end)
                        continue;
                    uint32_t* const kdst = s.dst.key[d];
                    uint32_t* const vdst = s.dst.val[d];
                    const uint32_t shift32 = (uint32_t(reinterpret_cast<uintptr_t>(kdst) >> 2) + begin) & 31u;
                    // slot p goes to kdst[p]; windows start at begin - shift32 (mod 2^32: the lanes before `begin` idle)
                    for (uint32_t p = begin - shift32 + warp * 32u + lane; int32_t(p - end) < 0; p += uint32_t(WARPS) * 32u)
                        if (int32_t(p - begin) >= 0)
                        {
                            kdst[p] = s.keys[p];
                            vdst[p] = s.vals[p];
                        }
                }
            }
            else if (full)
            {
#pragma unroll
                for (int k = 0; k < IPT; k++)
                {
                    const uint32_t p = tid + k * THREADS;
                    const uint32_t kk = s.keys[p];
                    if constexpr (PEER)
                    {
                        const uint32_t vv = s.vals[p];
                        const uint32_t d = digit_of(kk);
                        s.dst.key[d][p] = kk;
                        s.dst.val[d][p] = vv;
                    }
                    else if constexpr (KEYS_ONLY)
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
                    const uint32_t kk = s.keys[p];
                    if constexpr (PEER)
                    {
                        const uint32_t d = digit_of(kk);
                        s.dst.key[d][p] = kk;
                        s.dst.val[d][p] = s.vals[p];
                    }
                    else
                    {
                        const uint32_t dst = s.gbase[digit_of(kk)] + p;
                        keys_out[dst] = kk;
                        if constexpr (!KEYS_ONLY)
                            vals_out[dst] = s.vals[p];
                    }
                }
            }
        }

#include "glu_onesweep_ring.cuh"

        // ------------------------------------------------------------------------------------ small inputs
        //
        // Up to k_small_tile pairs: ONE CTA, ONE launch — the pairs stay in shared memory for all digit passes (count,
        // 256-digit scan, ballot ranking against per-warp running offsets, scatter, exactly the tile-local part of
        // onesweep_kernel).  The general path costs a memset + a histogram + a launch per digit with a chain hand-off
        // in each: 40-49 us whatever the size (profiles/r02_glu_test_benchmark.md); this one is a single short kernel.
        constexpr int k_small_threads = 256, k_small_ipt = 8, k_small_tile = k_small_threads * k_small_ipt;

        __global__ void __launch_bounds__(k_small_threads, 1)
            small_sort_kernel(uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t n, PassPlan plan)
        {
            constexpr int WARPS = k_small_threads / 32;
            __shared__ uint32_t s_keys[k_small_tile], s_vals[k_small_tile];
            __shared__ uint32_t s_hist[WARPS][k_radix];
            __shared__ uint32_t s_scan[k_radix / 32];
            const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            // slots past the end hold the largest key and sit behind every real pair: they rank after all real pairs
            // in every pass (equal digits keep their order) and are never written back
            for (uint32_t i = tid; i < uint32_t(k_small_tile); i += k_small_threads)
            {
                s_keys[i] = i < n ? keys[i] : 0xffffffffu;
                s_vals[i] = i < n ? vals[i] : 0u;
            }
            const uint32_t my_off = warp * (k_small_ipt * 32) + lane; // + i * 32 (warp-striped: input order)
            const uint32_t lt = lanemask_lt();
            uint32_t* wh = s_hist[warp];
            for (int p = 0; p < plan.num_passes; p++)
            {
                for (int i = tid; i < WARPS * k_radix; i += k_small_threads)
                    (&s_hist[0][0])[i] = 0u;
                __syncthreads(); // staged pairs (first pass), cleared counters
                uint32_t key[k_small_ipt], val[k_small_ipt], dig[k_small_ipt];
#pragma unroll
                for (int i = 0; i < k_small_ipt; i++)
                {
                    key[i] = s_keys[my_off + i * 32];
                    val[i] = s_vals[my_off + i * 32];
                    dig[i] = (key[i] >> plan.shift[p]) & plan.mask[p];
                    atomicAdd(&wh[dig[i]], 1u);
                }
                __syncthreads(); // every pair is in registers; the counts are final
                uint32_t total = 0, inc = 0;
                if (tid < k_radix)
                {
#pragma unroll
                    for (int w = 0; w < WARPS; w++)
                        total += s_hist[w][tid];
                    inc = total;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
                        if (lane >= unsigned(o))
                            inc += t;
                    }
                    if (lane == 31)
                        s_scan[warp] = inc;
                }
                __syncthreads();
                if (tid < k_radix)
                {
                    uint32_t running = inc - total;
                    for (unsigned w = 0; w < warp; w++)
                        running += s_scan[w];
#pragma unroll
                    for (int w = 0; w < WARPS; w++)
                    {
                        const uint32_t c = s_hist[w][tid];
                        s_hist[w][tid] = running; // first slot of (warp w, digit tid)
                        running += c;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < k_small_ipt; i++)
                {
                    const uint32_t peers = match_digit<Rank_Ballot>(dig[i]);
                    const uint32_t before = wh[dig[i]];
                    __syncwarp();
                    wh[dig[i]] = before + __popc(peers); // every peer stores the same value
                    __syncwarp();
                    const uint32_t r = before + __popc(peers & lt);
                    s_keys[r] = key[i];
                    s_vals[r] = val[i];
                }
                __syncthreads(); // the scatter is complete and nobody needs its running offsets any more
            }
            for (uint32_t i = tid; i < n; i += k_small_threads)
            {
                keys[i] = s_keys[i];
                vals[i] = s_vals[i];
            }
        }

        // ------------------------------------------------------------------------------------ host side

        struct SweepConfig
        {
            int id;
            int threads, ipt;
            int ring = 0; // 1: onesweep_ring_kernel (persistent CTAs, two-deep key ring), 2: the same with RATOM
        };
        constexpr SweepConfig k_configs[] = {
            // {id, threads, keys per thread}: tile = threads * ipt
            {0, 512, 16}, // 8192-pair tiles, 2 CTAs/SM
            {1, 384, 18}, // 6912, 3 CTAs/SM
            {2, 256, 16}, // 4096, 4 CTAs/SM (mid-size inputs)
            {3, 384, 20}, // 7680, 3 CTAs/SM
            {4, 512, 22}, // 11264, 2 CTAs/SM
            {5, 256, 8},  // 2048 (small inputs: more CTAs)
            {6, 384, 16}, // 6144, 3 CTAs/SM
            {7, 320, 18}, // 5760, 4 CTAs/SM
            {8, 320, 24}, // 7680, 3 CTAs/SM with 64 registers per thread
            // persistent ring kernel (glu_onesweep_ring.cuh): 12 B of shared memory per pair + the counters
            {9, 320, 16, 1},  // 5120, 3 CTAs/SM
            {10, 480, 16, 1}, // 7680, 2 CTAs/SM
            {11, 512, 14, 1}, // 7168, 2 CTAs/SM
            {12, 384, 20, 1}, // 7680, 2 CTAs/SM
            {13, 416, 18, 1}, // 7488, 2 CTAs/SM
            {14, 320, 16, 2}, // as 9..13 with the returning-atomic ranking loop
            {15, 480, 16, 2},
            {16, 512, 14, 2},
            {17, 384, 20, 2},
            {18, 416, 18, 2},
            // (4 CTAs/SM at 64 registers — 256 x 22 and 256 x 20 — measured 1.17 / 1.26 ms per pass against 1.09 ms:
            //  the per-tile fixed work, 256-digit scan + chain rows, outweighs the extra resident CTA)
        };
        constexpr int k_num_configs = int(sizeof(k_configs) / sizeof(k_configs[0]));

        int env_int(const char* name, int fallback)
        {
            const char* v = std::getenv(name);
            return v && *v ? std::atoi(v) : fallback;
        }

        // GLU_SORT_CHAIN_ROWS: rows per lane of a chain batch (2, 4, 8), +100 = 64 digits per chain CTA (4 chain CTAs
        // instead of 8).  Anything else would pair a chain shape with the wrong number of chain CTAs: default.
        int chain_rows_env()
        {
            static const int v = env_int("GLU_SORT_CHAIN_ROWS", 8);
            return (v == 2 || v == 4 || v == 8 || v == 104 || v == 108) ? v : 8;
        }

        // allow_forced = false: the flavoured kernels (glu_radix_sort_u32_ex) exist for the default shapes only
        const SweepConfig& select_config(size_t count, bool allow_forced = true)
        {
            static const int forced = env_int("GLU_SORT_CONFIG", -1); // tuning sweeps only
            if (allow_forced && forced >= 0 && forced < k_num_configs)
                return k_configs[forced];
            if (count <= (size_t(1) << 18))
                return k_configs[5];
            if (count <= (size_t(1) << 21))
                return k_configs[2];
            return k_configs[8];
        }

        bool use_tma_env()
        {
            static const int v = env_int("GLU_SORT_TMA", 1); // 0: force the ld/st staging path (tuning, tests)
            return v != 0;
        }

        int rank_mode()
        {
            static const int mode = env_int("GLU_SORT_RANK", Rank_Ballot); // match.any is ~1.6x slower on B200
            return mode == Rank_Match ? Rank_Match : Rank_Ballot;
        }

        struct TmpLayout
        {
            size_t tiles;
            size_t control_bytes; // tickets + histograms + look-back words: zeroed at the start of a sort
            size_t off_hist, off_lookback, off_keys, off_vals, total;
        };

        TmpLayout make_layout(size_t count, bool with_values = true, bool allow_forced = true)
        {
            const SweepConfig& c = select_config(count, allow_forced);
            TmpLayout l;
            const size_t tile = size_t(c.threads) * c.ipt;
            l.tiles = (count + tile - 1) / tile;
            l.off_hist = k_tmp_align; // tickets live in [0, 256)
            l.off_lookback = l.off_hist + k_max_passes * k_radix * sizeof(uint32_t);
            l.control_bytes = align_up(l.off_lookback + 2 * k_max_passes * l.tiles * k_radix * sizeof(uint32_t), k_tmp_align);
            l.off_keys = l.control_bytes;
            l.off_vals = l.off_keys + align_up(count * sizeof(uint32_t), k_tmp_align);
            l.total = l.off_vals + (with_values ? align_up(count * sizeof(uint32_t), k_tmp_align) : 0);
            return l;
        }

        template<int THREADS, int IPT, int MIN_BLOCKS, int MODE, bool PEER = false, bool DEST = false, int FLAVOR = 0,
                 bool SEG = false>
        int launch_sweep(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint32_t n, uint32_t shift,
                         uint32_t mask, const uint32_t* digit_offset, uint32_t* lookback, uint32_t* ticket,
                         unsigned tiles, cudaStream_t s, const uint32_t* d_n = nullptr,
                         uint32_t* const* key_dst = nullptr, uint32_t* const* val_dst = nullptr,
                         const uint8_t* dest_lut = nullptr, const uint2* tile_info = nullptr,
                         const uint32_t* tile_map = nullptr)
        {
            // per pass: `tiles` count rows followed by `tiles` prefix rows; grid = tiles + the chain CTAs
            uint32_t* prefix = lookback + size_t(tiles) * k_radix;
            auto kernel = onesweep_kernel<THREADS, IPT, MIN_BLOCKS, MODE, PEER, DEST, FLAVOR, SEG>;
            // TMA bulk copies need 16-byte aligned sources (tiles are multiples of 4 elements); vi is null without values
            const int allow_tma =
                ((reinterpret_cast<uintptr_t>(ki) | reinterpret_cast<uintptr_t>(vi)) & 15) == 0 && use_tma_env();
            constexpr size_t smem = sizeof(SweepSmem<THREADS, IPT, PEER, (FLAVOR & k_flavor_keys_only) == 0>);
            const int chain_rows = chain_rows_env();
            // bit 0: skip the look-back (timing experiments, wrong results); bit 1: tile ids from an atomic ticket
            // bits 8..: L2 prefetch distance in tiles (GLU_SORT_PREFETCH; 0 = off).  Default: one tile per SM ahead —
            // a third of the resident wave at 3 CTAs per SM; 74..444 measured within 1 % of each other at 2^28,
            // 888 (two waves: the lines are evicted again before they are used) 6 % slower than no prefetch.
            static const int prefetch_env = env_int("GLU_SORT_PREFETCH", -1);
            const int prefetch_tiles = prefetch_env >= 0 ? (prefetch_env > 0xffff ? 0xffff : prefetch_env) : current_sm_count();
            static const int options_env = env_int("GLU_SORT_OPTIONS", 0) & 0xff;
            const int options = options_env | (prefetch_tiles << 8);
            static std::atomic<bool> configured[64]; // per device; set after the attribute call (idempotent, so a race only repeats it)
            int dev = 0;
            GLU_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 64)
                return GLU_ERROR_INVALID_ARGUMENT;
            if (!configured[dev].load(std::memory_order_acquire))
            {
                GLU_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                configured[dev].store(true, std::memory_order_release);
            }
            const unsigned grid = tiles + (chain_rows >= 100 ? 4 : 8);
            ScopedKernelProfile prof(PEER ? GLU_KERNEL_SORT_PARTITION : GLU_KERNEL_SORT_ONESWEEP, s);
            kernel<<<grid, THREADS, smem, s>>>(ki, vi, ko, vo, n, shift, mask, digit_offset, lookback, prefix, ticket,
                                                    tiles, allow_tma, chain_rows, options, d_n, key_dst, val_dst, dest_lut,
                                                    tile_info, tile_map);
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }


        // onesweep_ring_kernel: a persistent grid — the chain CTAs plus as many tile CTAs as are resident at once
        template<int THREADS, int IPT, int MIN_BLOCKS, int MODE, int RATOM, int FLAVOR = 0, bool SEG = false>
        int launch_ring(const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo, uint32_t n, uint32_t shift,
                        uint32_t mask, const uint32_t* digit_offset, uint32_t* lookback, uint32_t* ticket, unsigned tiles,
                        cudaStream_t s, const uint32_t* d_n = nullptr, const uint2* tile_info = nullptr)
        {
            uint32_t* prefix = lookback + size_t(tiles) * k_radix;
            auto kernel = onesweep_ring_kernel<THREADS, IPT, MIN_BLOCKS, MODE, RATOM, FLAVOR, SEG>;
            const int allow_tma =
                ((reinterpret_cast<uintptr_t>(ki) | reinterpret_cast<uintptr_t>(vi)) & 15) == 0 && use_tma_env();
            constexpr size_t smem = sizeof(RingSmem<THREADS, IPT, (FLAVOR & k_flavor_keys_only) == 0>);
            const int chain_rows = chain_rows_env();
            // GLU_SORT_RING_TILES_PER_CTA: a CTA retires after that many tiles (0 = persistent until the tickets run out)
            static const int life = env_int("GLU_SORT_RING_TILES_PER_CTA", 0);
            static const int options = (env_int("GLU_SORT_OPTIONS", 0) & 0xff) | ((life > 0 ? (life < 3 ? 3 : life) : 0) << 8);
            static std::atomic<int> resident[64]; // CTAs per SM the hardware really grants, per device (0 = not asked yet)
            int dev = 0;
            GLU_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 64)
                return GLU_ERROR_INVALID_ARGUMENT;
            int per_sm = resident[dev].load(std::memory_order_acquire);
            if (per_sm == 0)
            {
                GLU_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                GLU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
                if (per_sm < 1)
                    return GLU_ERROR_CUDA;
                resident[dev].store(per_sm, std::memory_order_release);
            }
            const unsigned chain = chain_rows >= 100 ? 4 : 8;
            // GLU_SORT_RING_CTAS_PER_SM < occupancy leaves part of every SM to kernels of other streams (the multi-GPU
            // pipeline runs the next job's NVLink-bound exchange pass beside this sort)
            static const int share_env = env_int("GLU_SORT_RING_CTAS_PER_SM", 0);
            if (share_env > 0 && share_env < per_sm)
                per_sm = share_env;
            const unsigned capacity = unsigned(current_sm_count()) * unsigned(per_sm);
            // two tiles per ticket draw at start-up: no point in more CTAs than pairs of tiles
            unsigned workers = capacity > chain ? capacity - chain : 1;
            if (life > 0) // bounded CTA life: enough CTAs to draw every ticket, whatever is resident at once
                workers = (tiles + unsigned(life < 3 ? 3 : life) - 1) / unsigned(life < 3 ? 3 : life) + workers;
            if (workers > (tiles + 1) / 2)
                workers = (tiles + 1) / 2;
            ScopedKernelProfile prof(GLU_KERNEL_SORT_ONESWEEP, s);
            kernel<<<chain + workers, THREADS, smem, s>>>(ki, vi, ko, vo, n, shift, mask, digit_offset, lookback, prefix,
                                                          ticket, tiles, allow_tma, chain_rows, options, d_n, tile_info);
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }

        template<int MODE>
        int dispatch_sweep(const SweepConfig& c, const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo,
                           uint32_t n, uint32_t shift, uint32_t mask, const uint32_t* digit_offset, uint32_t* lookback,
                           uint32_t* ticket, unsigned tiles, cudaStream_t s, const uint32_t* d_n)
        {
            switch (c.id)
            {
#define GLU_SWEEP_CASE(ID, T, I, B)                                                                                    \
    case ID: return launch_sweep<T, I, B, MODE>(ki, vi, ko, vo, n, shift, mask, digit_offset, lookback, ticket, tiles, s, d_n);
                GLU_SWEEP_CASE(0, 512, 16, 2)
                GLU_SWEEP_CASE(1, 384, 18, 3)
                GLU_SWEEP_CASE(2, 256, 16, 4)
                GLU_SWEEP_CASE(3, 384, 20, 3)
                GLU_SWEEP_CASE(4, 512, 22, 2)
                GLU_SWEEP_CASE(6, 384, 16, 3)
                GLU_SWEEP_CASE(7, 320, 18, 4)
                GLU_SWEEP_CASE(8, 320, 24, 3)
#define GLU_RING_CASE(ID, T, I, B, A)                                                                                  \
    case ID:                                                                                                           \
        return launch_ring<T, I, B, MODE, A>(ki, vi, ko, vo, n, shift, mask, digit_offset, lookback, ticket, tiles, s, d_n);
                GLU_RING_CASE(9, 320, 16, 3, 0)
                GLU_RING_CASE(10, 480, 16, 2, 0)
                GLU_RING_CASE(11, 512, 14, 2, 0)
                GLU_RING_CASE(12, 384, 20, 2, 0)
                GLU_RING_CASE(13, 416, 18, 2, 0)
                GLU_RING_CASE(14, 320, 16, 3, 1)
                GLU_RING_CASE(15, 480, 16, 2, 1)
                GLU_RING_CASE(16, 512, 14, 2, 1)
                GLU_RING_CASE(17, 384, 20, 2, 1)
                GLU_RING_CASE(18, 416, 18, 2, 1)
#undef GLU_RING_CASE
            default: return launch_sweep<256, 8, 4, MODE>(ki, vi, ko, vo, n, shift, mask, digit_offset, lookback, ticket, tiles, s, d_n);
#undef GLU_SWEEP_CASE
            }
        }

        // glu_radix_sort_u32_ex: the three default tile shapes (select_config without the tuning override), ballot ranking
        template<int FLAVOR>
        int dispatch_sweep_flavor(const SweepConfig& c, const uint32_t* ki, const uint32_t* vi, uint32_t* ko, uint32_t* vo,
                                  uint32_t n, uint32_t shift, uint32_t mask, const uint32_t* digit_offset,
                                  uint32_t* lookback, uint32_t* ticket, unsigned tiles, cudaStream_t s)
        {
            switch (c.id)
            {
            case 8:
                return launch_sweep<320, 24, 3, Rank_Ballot, false, false, FLAVOR>(ki, vi, ko, vo, n, shift, mask, digit_offset,
                                                                                   lookback, ticket, tiles, s);
            case 2:
                return launch_sweep<256, 16, 4, Rank_Ballot, false, false, FLAVOR>(ki, vi, ko, vo, n, shift, mask, digit_offset,
                                                                                   lookback, ticket, tiles, s);
            case 5:
                return launch_sweep<256, 8, 4, Rank_Ballot, false, false, FLAVOR>(ki, vi, ko, vo, n, shift, mask, digit_offset,
                                                                                  lookback, ticket, tiles, s);
            default: return GLU_ERROR_INVALID_ARGUMENT;
            }
        }
    } // namespace
} // namespace glu_b200

using namespace glu_b200;

namespace
{
    // the partition pass always uses the large-input tile shape
    constexpr int k_part_threads = 320, k_part_ipt = 24, k_part_blocks = 3;

    // n: the count, or its upper bound when d_n (device-resident count) is given
    int launch_histogram(const uint32_t* d_keys, uint32_t n, int num_passes, uint32_t pre_shift, uint32_t key_mask,
                         uint32_t* hist, uint32_t* ticket, int make_offsets, int sms, cudaStream_t s,
                         const uint32_t* d_n = nullptr, int descending = 0)
    {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(d_keys);
        const uint32_t head = uint32_t(((16 - (addr & 15)) & 15) / sizeof(uint32_t));
        const uint32_t n_units = (n - (head < n ? head : n)) / 4;
        const size_t per_block = size_t(k_hist_threads) * k_hist_unroll;
        size_t grid = (size_t(n_units) + per_block - 1) / per_block;
        const size_t cap = size_t(sms) * k_hist_blocks_per_sm;
        grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
        const size_t smem = size_t(num_passes) * k_radix * k_hist_copies * sizeof(uint32_t);
        static std::atomic<bool> configured[64]; // per device; set after the attribute call (idempotent, so a race only repeats it)
        int dev = 0;
        GLU_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 64)
            return GLU_ERROR_INVALID_ARGUMENT;
        if (!configured[dev].load(std::memory_order_acquire))
        {
            GLU_CUDA_TRY(cudaFuncSetAttribute(histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              int(k_max_passes * k_radix * k_hist_copies * sizeof(uint32_t))));
            configured[dev].store(true, std::memory_order_release);
        }
        ScopedKernelProfile prof(GLU_KERNEL_SORT_HISTOGRAM, s);
        histogram_kernel<<<unsigned(grid), k_hist_threads, smem, s>>>(d_keys, n, d_n, head, num_passes, pre_shift,
                                                                       key_mask, hist, ticket, make_offsets, descending);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }
} // namespace

extern "C" size_t glu_radix_sort_u32kv_tmp_bytes(size_t count)
{
    if (count > k_max_count)
        return 0;
    if (count <= 1)
        return k_tmp_align;
    return make_layout(count).total;
}

namespace
{
    // count: the number of pairs, or (d_n != nullptr) the bound the grids and the scratch are sized for while the
    // actual number is read from *d_n by the kernels
    // flavor: k_flavor_* bits (non-zero: glu_radix_sort_u32_ex, d_vals null when keys only)
    int sort_impl(uint32_t* d_keys, uint32_t* d_vals, size_t count, const uint32_t* d_n, const PassPlan& plan, int flavor,
                  void* d_tmp, size_t tmp_bytes, glu_stream_t stream);
}

extern "C" int glu_radix_sort_u32kv(uint32_t* d_keys, uint32_t* d_vals, size_t count, size_t num_steps, void* d_tmp,
                                    size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_keys || !d_vals) // glu/RadixSort.hpp:275-276
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count <= 1) // glu/RadixSort.hpp:278-279
        return GLU_SUCCESS;
    return sort_impl(d_keys, d_vals, count, nullptr, make_pass_plan(num_steps), 0, d_tmp, tmp_bytes, stream);
}

extern "C" int glu_radix_sort_u32kv_dyn(uint32_t* d_keys, uint32_t* d_vals, const uint32_t* d_count, size_t max_count,
                                        size_t num_steps, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_keys || !d_vals || !d_count)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_count) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    if (max_count <= 1)
        return GLU_SUCCESS;
    return sort_impl(d_keys, d_vals, max_count, d_count, make_pass_plan(num_steps), 0, d_tmp, tmp_bytes, stream);
}

extern "C" size_t glu_radix_sort_u32_ex_tmp_bytes(size_t count, int with_values)
{
    if (count > k_max_count)
        return 0;
    if (count <= 1)
        return k_tmp_align;
    return make_layout(count, with_values != 0, false).total;
}

extern "C" int glu_radix_sort_u32_ex(uint32_t* d_keys, uint32_t* d_vals, size_t count, unsigned begin_bit,
                                     unsigned end_bit, int descending, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_keys || begin_bit > end_bit || end_bit > 32)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count <= 1 || begin_bit == end_bit) // no key bit takes part: every pair is "equal", stable = unchanged
        return GLU_SUCCESS;
    const int flavor = (d_vals ? 0 : k_flavor_keys_only) | (descending ? k_flavor_descending : 0);
    return sort_impl(d_keys, d_vals, count, nullptr, make_bit_plan(begin_bit, int(end_bit - begin_bit)), flavor | 0x100,
                     d_tmp, tmp_bytes, stream);
}

namespace
{
int sort_impl(uint32_t* d_keys, uint32_t* d_vals, size_t count, const uint32_t* d_n, const PassPlan& plan, int flavor_arg,
              void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    // bit 8 of flavor_arg: called through glu_radix_sort_u32_ex (default tile shapes, ballot ranking, its own layout)
    const bool ex = (flavor_arg & 0x100) != 0;
    const int flavor = flavor_arg & 0xff;
    const bool with_values = (flavor & k_flavor_keys_only) == 0;
    if (count > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if ((reinterpret_cast<uintptr_t>(d_keys) | reinterpret_cast<uintptr_t>(d_vals)) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    const TmpLayout l = make_layout(count, with_values, !ex);
    if (!d_tmp || tmp_bytes < l.total)
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    const int sms = current_sm_count();
    if (sms <= 0)
        return GLU_ERROR_CUDA;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // small inputs: one CTA, one launch, no scratch (GLU_SORT_SMALL_MAX=0 sends them down the general path)
    static const int small_max = env_int("GLU_SORT_SMALL_MAX", k_small_tile);
    if (flavor == 0 && !d_n && count <= size_t(small_max < k_small_tile ? small_max : k_small_tile))
    {
        ScopedKernelProfile prof(GLU_KERNEL_SORT_ONESWEEP, s);
        small_sort_kernel<<<1, k_small_threads, 0, s>>>(d_keys, d_vals, uint32_t(count), plan);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }
    char* tmp = static_cast<char*>(d_tmp);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(tmp); // [0..3] sweep passes, [4] histogram
    uint32_t* hist = reinterpret_cast<uint32_t*>(tmp + l.off_hist);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(tmp + l.off_lookback);
    uint32_t* alt_keys = reinterpret_cast<uint32_t*>(tmp + l.off_keys);
    uint32_t* alt_vals = with_values ? reinterpret_cast<uint32_t*>(tmp + l.off_vals) : nullptr;
    const uint32_t n = uint32_t(count);

    const size_t used_control = l.off_lookback + 2 * size_t(plan.num_passes) * l.tiles * k_radix * sizeof(uint32_t);
    // one memset over tickets, histograms and look-back words (286 MB at 2^28 pairs, ~0.04 ms); zeroing the look-back
    // words from inside the histogram kernel instead was measured: the histogram grows by the same 0.04 ms
    GLU_CUDA_TRY(cudaMemsetAsync(tmp, 0, used_control, s));

    {
        const int rc = launch_histogram(d_keys, n, plan.num_passes, plan.begin_bit, plan.key_mask, hist, tickets + 4, 1, sms,
                                        s, d_n, (flavor & k_flavor_descending) ? 1 : 0);
        if (rc != GLU_SUCCESS)
            return rc;
    }

    const SweepConfig& cfg = select_config(count, !ex);
    const int mode = rank_mode();
    uint32_t* kbuf[2] = {d_keys, alt_keys};
    uint32_t* vbuf[2] = {d_vals, alt_vals};
    for (int p = 0; p < plan.num_passes; p++)
    {
        const uint32_t* ki = kbuf[p & 1];
        const uint32_t* vi = vbuf[p & 1];
        uint32_t* ko = kbuf[(p + 1) & 1];
        uint32_t* vo = vbuf[(p + 1) & 1];
        uint32_t* lb = lookback + 2 * size_t(p) * l.tiles * k_radix;
        if (ex)
        {
            int rc = GLU_ERROR_INVALID_ARGUMENT;
            switch (flavor)
            {
#define GLU_FLAVOR_CASE(F)                                                                                             \
    case F:                                                                                                            \
        rc = dispatch_sweep_flavor<F>(cfg, ki, vi, ko, vo, n, plan.shift[p], plan.mask[p], hist + p * k_radix, lb,     \
                                      tickets + p, unsigned(l.tiles), s);                                              \
        break;
                GLU_FLAVOR_CASE(0)
                GLU_FLAVOR_CASE(1)
                GLU_FLAVOR_CASE(2)
                GLU_FLAVOR_CASE(3)
#undef GLU_FLAVOR_CASE
            }
            if (rc != GLU_SUCCESS)
                return rc;
            continue;
        }
        int rc = mode == Rank_Ballot
                     ? dispatch_sweep<Rank_Ballot>(cfg, ki, vi, ko, vo, n, plan.shift[p], plan.mask[p],
                                                   hist + p * k_radix, lb, tickets + p, unsigned(l.tiles), s, d_n)
                     : dispatch_sweep<Rank_Match>(cfg, ki, vi, ko, vo, n, plan.shift[p], plan.mask[p],
                                                  hist + p * k_radix, lb, tickets + p, unsigned(l.tiles), s, d_n);
        if (rc != GLU_SUCCESS)
            return rc;
    }
    if (plan.num_passes & 1)
    {
        // an odd number of passes leaves the result in the scratch: bring it home (the reference would
        // leave it there, glu/RadixSort.hpp:315-329 — documented deviation)
        GLU_CUDA_TRY(cudaMemcpyAsync(d_keys, alt_keys, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        if (with_values)
            GLU_CUDA_TRY(cudaMemcpyAsync(d_vals, alt_vals, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return GLU_SUCCESS;
}
} // namespace

// ------------------------------------------------------------------------------ multi-GPU building blocks


extern "C" int glu_radix_histogram_u32(const uint32_t* d_keys, size_t count, unsigned shift, unsigned bits,
                                       uint32_t* d_hist, glu_stream_t stream)
{
    if (!d_keys || !d_hist || bits == 0 || bits > 8 || shift > 31)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if (reinterpret_cast<uintptr_t>(d_keys) % sizeof(uint32_t) != 0 || reinterpret_cast<uintptr_t>(d_hist) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    const int sms = current_sm_count();
    if (sms <= 0)
        return GLU_ERROR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GLU_CUDA_TRY(cudaMemsetAsync(d_hist, 0, k_radix * sizeof(uint32_t), s));
    if (count == 0)
        return GLU_SUCCESS;
    return launch_histogram(d_keys, uint32_t(count), 1, shift, (1u << bits) - 1u, d_hist, nullptr, 0, sms, s);
}

extern "C" size_t glu_radix_partition_u32kv_tmp_bytes(size_t count)
{
    if (count > k_max_count)
        return 0;
    const size_t tile = size_t(k_part_threads) * k_part_ipt;
    const size_t tiles = (count + tile - 1) / tile;
    return k_tmp_align + align_up(2 * tiles * k_radix * sizeof(uint32_t), k_tmp_align);
}

namespace
{
    int partition_impl(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, const uint32_t* d_n, unsigned shift,
                       unsigned bits, uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp,
                       size_t tmp_bytes, glu_stream_t stream);
}

extern "C" int glu_radix_partition_u32kv(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, unsigned shift,
                                         unsigned bits, uint32_t* const* d_key_dst, uint32_t* const* d_val_dst,
                                         void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    return partition_impl(d_keys, d_vals, count, nullptr, shift, bits, d_key_dst, d_val_dst, d_tmp, tmp_bytes, stream);
}

extern "C" int glu_radix_partition_u32kv_dyn(const uint32_t* d_keys, const uint32_t* d_vals, const uint32_t* d_count,
                                             size_t max_count, unsigned shift, unsigned bits, uint32_t* const* d_key_dst,
                                             uint32_t* const* d_val_dst, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_count)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_count) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    return partition_impl(d_keys, d_vals, max_count, d_count, shift, bits, d_key_dst, d_val_dst, d_tmp, tmp_bytes, stream);
}

namespace
{
int partition_impl(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, const uint32_t* d_n, unsigned shift,
                   unsigned bits, uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp, size_t tmp_bytes,
                   glu_stream_t stream)
{
    if (!d_keys || !d_vals || !d_key_dst || !d_val_dst || bits == 0 || bits > 8 || shift > 31)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count == 0)
        return GLU_SUCCESS;
    if (count > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if ((reinterpret_cast<uintptr_t>(d_keys) | reinterpret_cast<uintptr_t>(d_vals)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst)) % sizeof(void*) != 0)
        return GLU_ERROR_MISALIGNED;
    const size_t need = glu_radix_partition_u32kv_tmp_bytes(count);
    if (!d_tmp || tmp_bytes < need)
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    if (current_sm_count() <= 0)
        return GLU_ERROR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t tile = size_t(k_part_threads) * k_part_ipt;
    const unsigned tiles = unsigned((count + tile - 1) / tile);
    char* tmp = static_cast<char*>(d_tmp);
    GLU_CUDA_TRY(cudaMemsetAsync(tmp, 0, need, s));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(tmp);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(tmp + k_tmp_align);
    return launch_sweep<k_part_threads, k_part_ipt, k_part_blocks, Rank_Ballot, true>(
        d_keys, d_vals, nullptr, nullptr, uint32_t(count), shift, (1u << bits) - 1u, nullptr, lookback, ticket, tiles, s,
        d_n, d_key_dst, d_val_dst);
}
} // namespace

namespace
{
    int partition_by_dest_impl(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, const uint32_t* d_n,
                               unsigned shift, unsigned bits, const uint8_t* d_dest_of_digit, uint32_t* const* d_key_dst,
                               uint32_t* const* d_val_dst, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);
}

extern "C" int glu_radix_partition_by_dest_u32kv(const uint32_t* d_keys, const uint32_t* d_vals, size_t count,
                                                 unsigned shift, unsigned bits, const uint8_t* d_dest_of_digit,
                                                 uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp,
                                                 size_t tmp_bytes, glu_stream_t stream)
{
    return partition_by_dest_impl(d_keys, d_vals, count, nullptr, shift, bits, d_dest_of_digit, d_key_dst, d_val_dst,
                                  d_tmp, tmp_bytes, stream);
}

extern "C" int glu_radix_partition_by_dest_u32kv_dyn(const uint32_t* d_keys, const uint32_t* d_vals,
                                                     const uint32_t* d_count, size_t max_count, unsigned shift,
                                                     unsigned bits, const uint8_t* d_dest_of_digit,
                                                     uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp,
                                                     size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_count)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_count) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    return partition_by_dest_impl(d_keys, d_vals, max_count, d_count, shift, bits, d_dest_of_digit, d_key_dst, d_val_dst,
                                  d_tmp, tmp_bytes, stream);
}

namespace
{
int partition_by_dest_impl(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, const uint32_t* d_n,
                           unsigned shift, unsigned bits, const uint8_t* d_dest_of_digit, uint32_t* const* d_key_dst,
                           uint32_t* const* d_val_dst, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_keys || !d_vals || !d_dest_of_digit || !d_key_dst || !d_val_dst || bits == 0 || bits > 8 || shift > 31)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count == 0)
        return GLU_SUCCESS;
    if (count > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if ((reinterpret_cast<uintptr_t>(d_keys) | reinterpret_cast<uintptr_t>(d_vals)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst)) % sizeof(void*) != 0)
        return GLU_ERROR_MISALIGNED;
    const size_t need = glu_radix_partition_u32kv_tmp_bytes(count);
    if (!d_tmp || tmp_bytes < need)
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    if (current_sm_count() <= 0)
        return GLU_ERROR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t tile = size_t(k_part_threads) * k_part_ipt;
    const unsigned tiles = unsigned((count + tile - 1) / tile);
    char* tmp = static_cast<char*>(d_tmp);
    GLU_CUDA_TRY(cudaMemsetAsync(tmp, 0, need, s));
    uint32_t* ticket = reinterpret_cast<uint32_t*>(tmp);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(tmp + k_tmp_align);
    return launch_sweep<k_part_threads, k_part_ipt, k_part_blocks, Rank_Ballot, true, true>(
        d_keys, d_vals, nullptr, nullptr, uint32_t(count), shift, (1u << bits) - 1u, nullptr, lookback, ticket, tiles, s,
        d_n, d_key_dst, d_val_dst, d_dest_of_digit);
}
} // namespace

// ------------------------------------------------------------------------------ exchange plan on the device

namespace glu_b200
{
namespace
{
    constexpr int k_max_plan_world = 16; // the by-destination partition handles < 16 destinations... 16 pointers

    // One CTA of 256 threads, thread b = bucket b.  The device-side twin of distributed.plan_exchange /
    // assign_buckets (gl-radix-sort_b200/distributed.py): same integer arithmetic, same results.
    __global__ void __launch_bounds__(k_radix, 1)
        exchange_plan_kernel(const uint32_t* __restrict__ hist_all, int world, int rank, uint32_t send_count,
                             uint64_t capacity, const uint64_t* __restrict__ peer_keys,
                             const uint64_t* __restrict__ peer_vals, uint64_t* key_dst, uint64_t* val_dst,
                             uint8_t* dest_of_digit, uint32_t* counts, uint64_t* info)
    {
        __shared__ unsigned long long s_warp[k_radix / 32];
        __shared__ uint32_t s_send[k_max_plan_world][k_max_plan_world]; // pairs source s sends to destination g
        __shared__ unsigned long long s_recv[k_max_plan_world];
        __shared__ int s_overflow;
        const unsigned b = threadIdx.x, lane = b & 31, warp = b >> 5;
        if (b < k_max_plan_world * k_max_plan_world)
            (&s_send[0][0])[b] = 0;
        if (b == 0)
            s_overflow = 0;
        uint32_t cnt[k_max_plan_world];
        unsigned long long bucket = 0;
#pragma unroll
        for (int src = 0; src < k_max_plan_world; src++)
        {
            cnt[src] = src < world ? hist_all[src * k_radix + b] : 0u;
            bucket += cnt[src];
        }
        // exclusive prefix of the bucket totals over b, and the job's total
        unsigned long long inc = bucket;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned long long t = __shfl_up_sync(k_full_mask, inc, o);
            if (lane >= unsigned(o))
                inc += t;
        }
        if (lane == 31)
            s_warp[warp] = inc;
        __syncthreads();
        unsigned long long before = inc - bucket, total = 0;
        for (unsigned w = 0; w < k_radix / 32; w++)
        {
            if (w < warp)
                before += s_warp[w];
            total += s_warp[w];
        }
        // a bucket goes to the rank in whose share of the global order its midpoint falls (assign_buckets)
        unsigned dest = 0;
        if (total)
        {
            const unsigned long long d = (2 * before + bucket) * (unsigned long long)world / (2 * total);
            dest = d < (unsigned long long)(world - 1) ? unsigned(d) : unsigned(world - 1);
        }
        dest_of_digit[b] = uint8_t(dest);
#pragma unroll
        for (int src = 0; src < k_max_plan_world; src++)
            if (src < world && cnt[src])
                atomicAdd(&s_send[src][dest], cnt[src]);
        __syncthreads();
        // destination g's receive buffer: the sources one after the other in rank order (source-major, stable)
        unsigned long long recv = 0, offset = 0;
        if (b < unsigned(world))
        {
            for (int src = 0; src < world; src++)
            {
                if (src < rank)
                    offset += s_send[src][b];
                recv += s_send[src][b];
            }
            s_recv[b] = recv;
            if (recv > capacity)
                atomicOr(&s_overflow, 1);
        }
        __syncthreads();
        const bool overflow = s_overflow != 0;
        key_dst[b] = (b < unsigned(world) && !overflow) ? peer_keys[b] + 4ull * offset : 0ull;
        val_dst[b] = (b < unsigned(world) && !overflow) ? peer_vals[b] + 4ull * offset : 0ull;
        if (b < unsigned(world))
            info[b] = recv;
        if (b == 0)
        {
            // on overflow nothing is sent and nothing is sorted; the host raises once it has read info[]
            counts[0] = overflow ? 0u : send_count;
            counts[1] = overflow ? 0u : uint32_t(s_recv[rank]);
            info[world] = s_recv[rank];
            info[world + 1] = overflow ? 1ull : 0ull;
        }
    }
} // namespace
} // namespace glu_b200

extern "C" int glu_radix_exchange_plan(const uint32_t* d_hist_all, int world, int rank, size_t send_count,
                                       size_t capacity, const uint64_t* d_peer_keys, const uint64_t* d_peer_vals,
                                       uint32_t** d_key_dst, uint32_t** d_val_dst, uint8_t* d_dest_of_digit,
                                       uint32_t* d_counts, uint64_t* d_info, glu_stream_t stream)
{
    if (!d_hist_all || !d_peer_keys || !d_peer_vals || !d_key_dst || !d_val_dst || !d_dest_of_digit || !d_counts ||
        !d_info || world < 1 || world > k_max_plan_world || rank < 0 || rank >= world)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (send_count > k_max_count || capacity > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if ((reinterpret_cast<uintptr_t>(d_hist_all) | reinterpret_cast<uintptr_t>(d_counts)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_peer_keys) | reinterpret_cast<uintptr_t>(d_peer_vals) |
         reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst) |
         reinterpret_cast<uintptr_t>(d_info)) % sizeof(uint64_t) != 0)
        return GLU_ERROR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    exchange_plan_kernel<<<1, k_radix, 0, s>>>(d_hist_all, world, rank, uint32_t(send_count), uint64_t(capacity),
                                               d_peer_keys, d_peer_vals, reinterpret_cast<uint64_t*>(d_key_dst),
                                               reinterpret_cast<uint64_t*>(d_val_dst), d_dest_of_digit, d_counts, d_info);
    GLU_LAUNCH_CHECK();
    return GLU_SUCCESS;
}

// ------------------------------------------------------------------------------ segmented sort (many sorts, one launch set)
#include "glu_radix_sort_seg.cuh"

// ------------------------------------------------------------------------------ bucket-major exchange plan (MSD split + segmented local sort)

namespace glu_b200
{
namespace
{
    // One CTA of 256 threads, thread b = bucket b.  Same bucket -> rank assignment as exchange_plan_kernel; the receive
    // layout is BUCKET-major: rank g's buckets in increasing order, each starting at a tile boundary of the segmented
    // sort (glu_radix_sort_segment_tile()), inside a bucket the sources in rank order (stable).
    __global__ void __launch_bounds__(k_radix, 1)
        exchange_plan_buckets_kernel(const uint32_t* __restrict__ hist_all, int world, int rank, uint32_t send_count,
                                     uint32_t capacity_tiles, uint32_t tile, const uint64_t* __restrict__ peer_keys,
                                     const uint64_t* __restrict__ peer_vals, uint64_t* key_dst, uint64_t* val_dst,
                                     uint32_t* seg_count, uint32_t* counts, uint64_t* info)
    {
        __shared__ unsigned long long s_warp[k_radix / 32];
        __shared__ uint32_t s_warp_tiles[k_radix / 32];
        __shared__ uint32_t s_first_tiles[k_max_plan_world]; // tiles in front of the first bucket of each destination
        __shared__ uint32_t s_recv_tiles[k_max_plan_world];
        __shared__ unsigned long long s_recv[k_max_plan_world];
        __shared__ int s_overflow;
        const unsigned b = threadIdx.x, lane = b & 31, warp = b >> 5;
        if (b < k_max_plan_world)
        {
            s_first_tiles[b] = 0xffffffffu;
            s_recv_tiles[b] = 0;
            s_recv[b] = 0;
        }
        if (b == 0)
            s_overflow = 0;
        unsigned long long bucket = 0, before_me = 0; // pairs of this bucket: all sources / sources < rank
#pragma unroll
        for (int src = 0; src < k_max_plan_world; src++)
        {
            const uint32_t c = src < world ? hist_all[src * k_radix + b] : 0u;
            bucket += c;
            if (src < rank)
                before_me += c;
        }
        const uint32_t tiles = uint32_t((bucket + tile - 1) / tile);
        unsigned long long inc = bucket;
        uint32_t inc_t = tiles;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned long long t = __shfl_up_sync(k_full_mask, inc, o);
            const uint32_t tt = __shfl_up_sync(k_full_mask, inc_t, o);
            if (lane >= unsigned(o))
            {
                inc += t;
                inc_t += tt;
            }
        }
        if (lane == 31)
        {
            s_warp[warp] = inc;
            s_warp_tiles[warp] = inc_t;
        }
        __syncthreads();
        unsigned long long before = inc - bucket, total = 0;
        uint32_t tiles_before = inc_t - tiles;
        for (unsigned w = 0; w < k_radix / 32; w++)
        {
            if (w < warp)
            {
                before += s_warp[w];
                tiles_before += s_warp_tiles[w];
            }
            total += s_warp[w];
        }
        unsigned dest = 0;
        if (total)
        {
            const unsigned long long d = (2 * before + bucket) * (unsigned long long)world / (2 * total);
            dest = d < (unsigned long long)(world - 1) ? unsigned(d) : unsigned(world - 1);
        }
        atomicMin(&s_first_tiles[dest], tiles_before);
        atomicAdd(&s_recv_tiles[dest], tiles);
        atomicAdd(&s_recv[dest], bucket);
        __syncthreads();
        if (b < unsigned(world) && s_recv_tiles[b] > capacity_tiles)
            atomicOr(&s_overflow, 1);
        __syncthreads();
        const bool overflow = s_overflow != 0;
        const unsigned long long offset = (unsigned long long)(tiles_before - s_first_tiles[dest]) * tile + before_me;
        key_dst[b] = overflow ? 0ull : peer_keys[dest] + 4ull * offset;
        val_dst[b] = overflow ? 0ull : peer_vals[dest] + 4ull * offset;
        seg_count[b] = (!overflow && dest == unsigned(rank)) ? uint32_t(bucket) : 0u;
        if (b < unsigned(world))
            info[b] = s_recv[b];
        if (b == 0)
        {
            counts[0] = overflow ? 0u : send_count;
            counts[1] = overflow ? 0u : uint32_t(s_recv[rank]);
            info[world] = s_recv[rank];
            info[world + 1] = overflow ? 1ull : 0ull;
        }
    }
} // namespace
} // namespace glu_b200

extern "C" int glu_radix_exchange_plan_buckets(const uint32_t* d_hist_all, int world, int rank, size_t send_count,
                                               size_t capacity_tiles, const uint64_t* d_peer_keys,
                                               const uint64_t* d_peer_vals, uint32_t** d_key_dst, uint32_t** d_val_dst,
                                               uint32_t* d_seg_count, uint32_t* d_counts, uint64_t* d_info,
                                               glu_stream_t stream)
{
    if (!d_hist_all || !d_peer_keys || !d_peer_vals || !d_key_dst || !d_val_dst || !d_seg_count || !d_counts || !d_info ||
        world < 1 || world > k_max_plan_world || rank < 0 || rank >= world)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (send_count > k_max_count || capacity_tiles * glu_radix_sort_segment_tile() > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if ((reinterpret_cast<uintptr_t>(d_hist_all) | reinterpret_cast<uintptr_t>(d_counts) |
         reinterpret_cast<uintptr_t>(d_seg_count)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_peer_keys) | reinterpret_cast<uintptr_t>(d_peer_vals) |
         reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst) |
         reinterpret_cast<uintptr_t>(d_info)) % sizeof(uint64_t) != 0)
        return GLU_ERROR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    exchange_plan_buckets_kernel<<<1, k_radix, 0, s>>>(
        d_hist_all, world, rank, uint32_t(send_count), uint32_t(capacity_tiles), uint32_t(glu_radix_sort_segment_tile()),
        d_peer_keys, d_peer_vals, reinterpret_cast<uint64_t*>(d_key_dst), reinterpret_cast<uint64_t*>(d_val_dst), d_seg_count,
        d_counts, d_info);
    GLU_LAUNCH_CHECK();
    return GLU_SUCCESS;
}

// ------------------------------------------------------------------------------ staged exchange: local MSD pass + long-run peer copy

namespace glu_b200
{
namespace
{
    // One CTA of 256 threads, thread b = bucket b: where this rank's pairs of bucket b go in its own bucket-major
    // staging arrays (exclusive scan of its 256 counts) — the pointer table of the local MSD pass.
    // Buckets this rank keeps (seg_count[b] != 0) skip the staging arrays: the MSD pass writes them straight to their
    // place in the rank's own receive arrays (final_key_dst[b]) and the copy kernel gets a count of 0 for them.
    __global__ void __launch_bounds__(k_radix, 1)
        exchange_stage_tables_kernel(const uint32_t* __restrict__ my_hist, uint64_t stage_keys, uint64_t stage_vals,
                                     const uint32_t* __restrict__ seg_count, const uint64_t* __restrict__ final_key_dst,
                                     const uint64_t* __restrict__ final_val_dst, uint64_t* stage_key_dst,
                                     uint64_t* stage_val_dst, uint32_t* stage_off, uint32_t* copy_count)
    {
        __shared__ uint32_t s_warp[k_radix / 32];
        const unsigned b = threadIdx.x, lane = b & 31, warp = b >> 5;
        const uint32_t c = my_hist[b];
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
            if (lane >= unsigned(o))
                inc += t;
        }
        if (lane == 31)
            s_warp[warp] = inc;
        __syncthreads();
        uint32_t off = inc - c;
        for (unsigned w = 0; w < warp; w++)
            off += s_warp[w];
        const bool mine = seg_count && seg_count[b] != 0 && final_key_dst[b] != 0;
        stage_off[b] = off;
        stage_key_dst[b] = mine ? final_key_dst[b] : stage_keys + 4ull * off;
        stage_val_dst[b] = mine ? final_val_dst[b] : stage_vals + 4ull * off;
        copy_count[b] = mine ? 0u : c;
    }

    // The all-to-all itself: this rank's run of bucket b — count[b] pairs at stage_off[b] of the staging arrays — goes to
    // key_dst[b] / val_dst[b] (peer memory over NVLink for remote buckets).  Runs are long (count / 256 pairs), so every
    // warp-wide store is one full, aligned 128-byte line of its destination: the row grid of a bucket starts at the
    // line its destination starts in (only the first and last line of a run are partial).  Few CTAs saturate the links;
    // the rest of the GPU keeps sorting (the two-lane pipeline runs this under the previous job's local sort).
    constexpr int k_copy_threads = 256, k_copy_unroll = 4;
    __global__ void __launch_bounds__(k_copy_threads)
        exchange_copy_kernel(const uint32_t* __restrict__ stage_keys, const uint32_t* __restrict__ stage_vals,
                             const uint32_t* __restrict__ stage_off, const uint32_t* __restrict__ count,
                             uint32_t* const* __restrict__ key_dst, uint32_t* const* __restrict__ val_dst)
    {
        __shared__ uint32_t s_off[k_radix], s_cnt[k_radix], s_shift[k_radix], s_rows[k_radix + 1], s_warp[k_radix / 32];
        __shared__ uint32_t* s_kd[k_radix];
        __shared__ uint32_t* s_vd[k_radix];
        const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        {
            uint32_t* kd = key_dst[tid];
            uint32_t c = kd ? count[tid] : 0u; // a null destination: the plan found an overflow, nothing moves
            const uint32_t shift32 = uint32_t(reinterpret_cast<uintptr_t>(kd) >> 2) & 31u;
            const uint32_t rows = c ? (shift32 + c + 31u) / 32u : 0u;
            s_off[tid] = stage_off[tid];
            s_cnt[tid] = c;
            s_shift[tid] = shift32;
            s_kd[tid] = kd;
            s_vd[tid] = val_dst[tid];
            uint32_t inc = rows;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
                if (lane >= unsigned(o))
                    inc += t;
            }
            if (lane == 31)
                s_warp[warp] = inc;
            __syncthreads();
            uint32_t before = inc - rows;
            for (unsigned w = 0; w < warp; w++)
                before += s_warp[w];
            s_rows[tid] = before;
            if (tid == k_radix - 1)
                s_rows[k_radix] = before + rows;
            __syncthreads();
        }
        const uint32_t total_rows = s_rows[k_radix];
        const uint32_t warps = gridDim.x * (k_copy_threads / 32);
        const uint32_t per = (total_rows + warps - 1) / warps; // a contiguous range of rows per warp
        const uint32_t gw = blockIdx.x * (k_copy_threads / 32) + warp;
        uint32_t r = gw * per;
        const uint32_t r_end = r + per < total_rows ? r + per : total_rows;
        if (r >= r_end)
            return;
        // bucket of row r: the last b with s_rows[b] <= r among the non-empty ones
        uint32_t b;
        {
            uint32_t lo = 0, hi = k_radix;
            while (hi - lo > 1)
            {
                const uint32_t mid = (lo + hi) / 2;
                if (s_rows[mid] <= r)
                    lo = mid;
                else
                    hi = mid;
            }
            b = lo;
        }
        while (r < r_end)
        {
            uint32_t kk[k_copy_unroll], vv[k_copy_unroll];
            uint32_t* kp[k_copy_unroll];
            uint32_t* vp[k_copy_unroll];
#pragma unroll
            for (int u = 0; u < k_copy_unroll; u++)
            {
                kp[u] = nullptr;
                vp[u] = nullptr;
                if (r < r_end)
                {
                    while (r >= s_rows[b + 1]) // rows are consecutive: the bucket only moves forward
                        b++;
                    const int32_t e = int32_t((r - s_rows[b]) * 32u + lane) - int32_t(s_shift[b]);
                    if (e >= 0 && uint32_t(e) < s_cnt[b])
                    {
                        kk[u] = ld_stream_u32(stage_keys + s_off[b] + e);
                        vv[u] = ld_stream_u32(stage_vals + s_off[b] + e);
                        kp[u] = s_kd[b] + e;
                        vp[u] = s_vd[b] + e;
                    }
                    r++;
                }
            }
#pragma unroll
            for (int u = 0; u < k_copy_unroll; u++)
                if (kp[u])
                {
                    *kp[u] = kk[u];
                    *vp[u] = vv[u];
                }
        }
    }
} // namespace
} // namespace glu_b200

extern "C" int glu_radix_exchange_stage_tables(const uint32_t* d_my_hist, uint32_t* d_stage_keys, uint32_t* d_stage_vals,
                                               const uint32_t* d_seg_count, uint32_t* const* d_key_dst,
                                               uint32_t* const* d_val_dst, uint32_t** d_stage_key_dst,
                                               uint32_t** d_stage_val_dst, uint32_t* d_stage_off, uint32_t* d_copy_count,
                                               glu_stream_t stream)
{
    if (!d_my_hist || !d_stage_keys || !d_stage_vals || !d_stage_key_dst || !d_stage_val_dst || !d_stage_off ||
        !d_copy_count || (d_seg_count && (!d_key_dst || !d_val_dst)))
        return GLU_ERROR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_my_hist) | reinterpret_cast<uintptr_t>(d_stage_off) |
         reinterpret_cast<uintptr_t>(d_stage_keys) | reinterpret_cast<uintptr_t>(d_stage_vals) |
         reinterpret_cast<uintptr_t>(d_seg_count) | reinterpret_cast<uintptr_t>(d_copy_count)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_stage_key_dst) | reinterpret_cast<uintptr_t>(d_stage_val_dst) |
         reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst)) % sizeof(uint64_t) != 0)
        return GLU_ERROR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    exchange_stage_tables_kernel<<<1, k_radix, 0, s>>>(
        d_my_hist, reinterpret_cast<uint64_t>(d_stage_keys), reinterpret_cast<uint64_t>(d_stage_vals), d_seg_count,
        reinterpret_cast<const uint64_t*>(d_key_dst), reinterpret_cast<const uint64_t*>(d_val_dst),
        reinterpret_cast<uint64_t*>(d_stage_key_dst), reinterpret_cast<uint64_t*>(d_stage_val_dst), d_stage_off,
        d_copy_count);
    GLU_LAUNCH_CHECK();
    return GLU_SUCCESS;
}

extern "C" int glu_radix_exchange_copy_u32kv(const uint32_t* d_stage_keys, const uint32_t* d_stage_vals,
                                             const uint32_t* d_stage_off, const uint32_t* d_count,
                                             uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, int num_ctas,
                                             glu_stream_t stream)
{
    if (!d_stage_keys || !d_stage_vals || !d_stage_off || !d_count || !d_key_dst || !d_val_dst)
        return GLU_ERROR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(d_stage_keys) | reinterpret_cast<uintptr_t>(d_stage_vals) |
         reinterpret_cast<uintptr_t>(d_stage_off) | reinterpret_cast<uintptr_t>(d_count)) % sizeof(uint32_t) != 0 ||
        (reinterpret_cast<uintptr_t>(d_key_dst) | reinterpret_cast<uintptr_t>(d_val_dst)) % sizeof(void*) != 0)
        return GLU_ERROR_MISALIGNED;
    const int sms = current_sm_count();
    if (sms <= 0)
        return GLU_ERROR_CUDA;
    // default: one CTA per SM (GLU_EXCHANGE_COPY_CTAS for sweeps) — NVLink-bound, it needs no more
    static const int env_ctas = env_int("GLU_EXCHANGE_COPY_CTAS", 0);
    const int ctas = num_ctas > 0 ? num_ctas : (env_ctas > 0 ? env_ctas : sms);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ScopedKernelProfile prof(GLU_KERNEL_SORT_PARTITION, s);
    exchange_copy_kernel<<<unsigned(ctas), k_copy_threads, 0, s>>>(d_stage_keys, d_stage_vals, d_stage_off, d_count, d_key_dst,
                                                                   d_val_dst);
    GLU_LAUNCH_CHECK();
    return GLU_SUCCESS;
}
