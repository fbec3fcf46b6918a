// glu_scan.cu — glu_scan_exclusive(): the B200 replacement for glu::BlellochScan::operator()
// (glu/BlellochScan.hpp:130-139, host loops :142-190, shaders :13-76).
//
// The reference runs a Blelloch up-sweep and down-sweep in GLOBAL memory, one dispatch per tree level
// (2*log2(N) dispatches, stride-2^k gathers, ~5 element-sized transfers per element, power-of-two N
// only).  Here the scan is ONE kernel launch that reads every element once and writes it once
// (8 B of HBM traffic per 32-bit element):
//   * a tile of THREADS x VPT 16-byte vectors is loaded warp-striped with 128-bit streaming loads,
//     scanned in registers (4-element serial scan + shuffle scan across the warp + chunk chaining);
//   * tiles are chained with a decoupled look-back (Merrill & Garland): each tile publishes its
//     aggregate, then walks back over its predecessors' {aggregate | inclusive prefix} words, a warp
//     at a time, until it meets an inclusive prefix;  status and value share one 64-bit word, so a
//     single relaxed 64-bit store / load is the whole protocol for 4-byte element types;
//   * tile ids come from an atomic ticket, so a tile's predecessors are always already running
//     (forward progress does not depend on the hardware's CTA scheduling order);
//   * `num_partitions` adjacent segments are native: tiles never straddle a segment, the first tile
//     of each segment publishes an inclusive prefix directly and look-back stops at the segment start;
//   * any `count` is accepted (the reference aborts on non powers of two, glu/BlellochScan.hpp:134).
// Element types wider than 4 bytes (double, vecN, dvecN, ...) use the same structure with one
// element per lane per item and a {flag, aggregate, inclusive} record published with release /
// acquire ordering.
#include <cstdlib>

#include "glu_common.cuh"

namespace glu_b200
{
    namespace
    {
        constexpr uint64_t k_flag_aggregate = 1ull << 32;
        constexpr uint64_t k_flag_inclusive = 2ull << 32;

        template<typename T> __device__ __forceinline__ uint32_t to_bits(T v);
        template<> __device__ __forceinline__ uint32_t to_bits<uint32_t>(uint32_t v) { return v; }
        template<> __device__ __forceinline__ uint32_t to_bits<float>(float v) { return __float_as_uint(v); }
        template<typename T> __device__ __forceinline__ T from_bits(uint32_t v);
        template<> __device__ __forceinline__ uint32_t from_bits<uint32_t>(uint32_t v) { return v; }
        template<> __device__ __forceinline__ float from_bits<float>(uint32_t v) { return __uint_as_float(v); }

        template<typename T> struct is_float_type
        {
            static constexpr bool value = false;
        };
        template<> struct is_float_type<float>
        {
            static constexpr bool value = true;
        };

        // inclusive scan of `v` across the warp
        template<typename T> __device__ __forceinline__ T warp_inclusive_scan(T v, unsigned lane)
        {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                T t = __shfl_up_sync(k_full_mask, v, o);
                if (lane >= unsigned(o))
                    v += t;
            }
            return v;
        }

        template<typename T> __device__ __forceinline__ T warp_sum(T v)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_xor_sync(k_full_mask, v, o);
            return v;
        }
        template<> __device__ __forceinline__ uint32_t warp_sum<uint32_t>(uint32_t v)
        {
            return __reduce_add_sync(k_full_mask, v);
        }

        // ------------------------------------------------------------------------ 4-byte element types
        //
        // T = uint32_t (also serves Int: two's-complement addition is the same bit pattern) or float.
        //
        // TICKET = true : tile id from an atomic ticket (forward progress guaranteed by construction);
        // TICKET = false: tile id = blockIdx.x (relies on the in-order CTA dispatch of the hardware, like
        //                 CUB's DeviceScan; saves one L2 round trip and one barrier per tile).
        template<typename T, int THREADS, int VPT, int MIN_BLOCKS, bool TICKET>
        __global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
            scan_b32_kernel(T* __restrict__ data, size_t count, uint32_t tiles_per_part, uint32_t* ticket,
                            uint64_t* state, int debug_no_lookback, const T* __restrict__ init)
        {
            constexpr int TILE = THREADS * VPT * 4;
            constexpr int WARPS = THREADS / 32;
            constexpr int WARP_ELEMS = VPT * 128;
            static_assert(WARPS <= 32, "one warp scans the warp totals");

            __shared__ T s_warp_total[WARPS];
            __shared__ T s_warp_prefix[WARPS];
            __shared__ T s_tile_prefix;
            __shared__ uint32_t s_tile;

            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

            uint32_t tile = blockIdx.x;
            if (TICKET)
            {
                if (threadIdx.x == 0)
                    s_tile = atomicAdd(ticket, 1u);
                __syncthreads();
                tile = s_tile;
            }
            const uint32_t part = tile / tiles_per_part;
            const uint32_t tp = tile - part * tiles_per_part; // tile index inside its partition
            const size_t in_part = size_t(tp) * TILE;
            const size_t base = size_t(part) * count + in_part;
            const uint32_t valid = uint32_t(count - in_part < size_t(TILE) ? count - in_part : size_t(TILE));
            const bool vec = valid == uint32_t(TILE) && (reinterpret_cast<uintptr_t>(data + base) & 15) == 0;
            const uint32_t my_off = warp * WARP_ELEMS + lane * 4; // + j * 128

            T x[VPT][4];
            if (vec)
            {
#pragma unroll
                for (int j = 0; j < VPT; j++)
                {
                    uint4 r = ld_stream_v4(data + base + my_off + j * 128);
                    x[j][0] = from_bits<T>(r.x);
                    x[j][1] = from_bits<T>(r.y);
                    x[j][2] = from_bits<T>(r.z);
                    x[j][3] = from_bits<T>(r.w);
                }
            }
            else
            {
#pragma unroll
                for (int j = 0; j < VPT; j++)
#pragma unroll
                    for (int c = 0; c < 4; c++)
                    {
                        uint32_t idx = my_off + j * 128 + c;
                        x[j][c] = idx < valid ? data[base + idx] : T(0);
                    }
            }

            // per-vector sums, then VPT independent warp scans, then chain the VPT chunks of the warp
            T sum4[VPT], inc[VPT], ex[VPT];
#pragma unroll
            for (int j = 0; j < VPT; j++)
                inc[j] = sum4[j] = x[j][0] + x[j][1] + x[j][2] + x[j][3];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                for (int j = 0; j < VPT; j++)
                {
                    T t = __shfl_up_sync(k_full_mask, inc[j], o);
                    if (lane >= unsigned(o))
                        inc[j] += t;
                }
            T chunk_base = T(0);
#pragma unroll
            for (int j = 0; j < VPT; j++)
            {
                T total = __shfl_sync(k_full_mask, inc[j], 31);
                T lane_ex;
                if (is_float_type<T>::value)
                {
                    lane_ex = __shfl_up_sync(k_full_mask, inc[j], 1); // exact: no inclusive-minus-own rounding
                    if (lane == 0)
                        lane_ex = T(0);
                }
                else
                    lane_ex = inc[j] - sum4[j];
                ex[j] = chunk_base + lane_ex;
                chunk_base += total;
            }
            if (lane == 0)
                s_warp_total[warp] = chunk_base;
            __syncthreads();

            if (warp == 0)
            {
                T wt = lane < WARPS ? s_warp_total[lane] : T(0);
                T winc = warp_inclusive_scan(wt, lane);
                T wex = __shfl_up_sync(k_full_mask, winc, 1);
                if (lane == 0)
                    wex = T(0);
                if (lane < WARPS)
                    s_warp_prefix[lane] = wex;
                const T aggregate = __shfl_sync(k_full_mask, winc, 31);

                // the first tile of a partition starts from `init` (std::exclusive_scan's init; 0 if absent)
                T exclusive = (tp == 0 && init) ? *init : T(0);
                if (tp == 0 || debug_no_lookback) // debug_no_lookback: timing experiments only (wrong results)
                {
                    if (lane == 0)
                        st_relaxed_u64(&state[tile], k_flag_inclusive | to_bits<T>(exclusive + aggregate));
                }
                else
                {
                    if (lane == 0)
                        st_relaxed_u64(&state[tile], k_flag_aggregate | to_bits<T>(aggregate));
                    // decoupled look-back, 32 predecessors per step; lanes past the partition start behave
                    // as an inclusive prefix of 0, which ends the walk
                    uint32_t remaining = tp;
                    uint32_t pred = tile - 1;
                    while (true)
                    {
                        const bool in_range = lane < remaining;
                        uint64_t w;
                        uint32_t inclusive_mask;
                        while (true)
                        {
                            w = in_range ? ld_relaxed_u64(&state[pred - lane]) : k_flag_inclusive;
                            const uint32_t flag = uint32_t(w >> 32);
                            const uint32_t empty_mask = __ballot_sync(k_full_mask, flag == 0);
                            inclusive_mask = __ballot_sync(k_full_mask, flag == 2);
                            // only the lanes up to the first inclusive prefix matter
                            const uint32_t need =
                                inclusive_mask ? ((2u << (__ffs(inclusive_mask) - 1)) - 1u) : k_full_mask;
                            if ((empty_mask & need) == 0)
                                break;
                        }
                        const uint32_t first = inclusive_mask ? uint32_t(__ffs(inclusive_mask) - 1) : 31u;
                        T contrib = lane <= first ? from_bits<T>(uint32_t(w)) : T(0);
                        exclusive += warp_sum(contrib);
                        if (inclusive_mask)
                            break;
                        pred -= 32;
                        remaining -= 32;
                    }
                    if (lane == 0)
                        st_relaxed_u64(&state[tile], k_flag_inclusive | to_bits<T>(exclusive + aggregate));
                }
                if (lane == 0)
                    s_tile_prefix = exclusive;
            }
            __syncthreads();

            const T prefix = s_tile_prefix + s_warp_prefix[warp];
            if (vec)
            {
#pragma unroll
                for (int j = 0; j < VPT; j++)
                {
                    T b = prefix + ex[j];
                    uint4 r;
                    r.x = to_bits<T>(b);
                    b += x[j][0];
                    r.y = to_bits<T>(b);
                    b += x[j][1];
                    r.z = to_bits<T>(b);
                    b += x[j][2];
                    r.w = to_bits<T>(b);
                    st_stream_v4(data + base + my_off + j * 128, r);
                }
            }
            else
            {
#pragma unroll
                for (int j = 0; j < VPT; j++)
                {
                    T b = prefix + ex[j];
#pragma unroll
                    for (int c = 0; c < 4; c++)
                    {
                        uint32_t idx = my_off + j * 128 + c;
                        if (idx < valid)
                            data[base + idx] = b;
                        b += x[j][c];
                    }
                }
            }
        }

        // Decoupled look-back over `state`, one warp: returns the sum of all earlier tiles of the
        // partition (tp of them), 32 predecessors per round trip; lanes past the partition start act as
        // an inclusive prefix of 0, which ends the walk.
        template<typename T>
        __device__ __forceinline__ T lookback_walk(const uint64_t* state, uint32_t tile, uint32_t tp, unsigned lane)
        {
            T exclusive = T(0);
            uint32_t remaining = tp;
            uint32_t pred = tile - 1;
            while (true)
            {
                const bool in_range = lane < remaining;
                uint64_t w;
                uint32_t inclusive_mask;
                while (true)
                {
                    w = in_range ? ld_relaxed_u64(&state[pred - lane]) : k_flag_inclusive;
                    const uint32_t flag = uint32_t(w >> 32);
                    const uint32_t empty_mask = __ballot_sync(k_full_mask, flag == 0);
                    inclusive_mask = __ballot_sync(k_full_mask, flag == 2);
                    // only the lanes up to the first inclusive prefix matter
                    const uint32_t need = inclusive_mask ? ((2u << (__ffs(inclusive_mask) - 1)) - 1u) : k_full_mask;
                    if ((empty_mask & need) == 0)
                        break;
                }
                const uint32_t first = inclusive_mask ? uint32_t(__ffs(inclusive_mask) - 1) : 31u;
                T contrib = lane <= first ? from_bits<T>(uint32_t(w)) : T(0);
                exclusive += warp_sum(contrib);
                if (inclusive_mask)
                    break;
                pred -= 32;
                remaining -= 32;
            }
            return exclusive;
        }

        // ------------------------------------------- 4-byte element types: persistent, TMA, chain warps
        //
        // Measured on B200 the one-tile-per-CTA kernel above runs at 6.8 TB/s when the look-back
        // dependency is cut (GLU_SCAN_DEBUG_NO_LOOKBACK) and at 3.6 TB/s with it: every CTA sits idle,
        // with no loads in flight, while it waits for its predecessors.  This kernel takes the chain off
        // the data path with warp specialisation inside resident CTAs:
        //   * a PRODUCER lane takes tile tickets and streams whole tiles into a ring of shared-memory
        //     stages with cp.async.bulk (TMA), completion on an mbarrier per stage;
        //   * AGGREGATOR warps (alternating tiles) reduce a stage the moment it lands and publish the tile
        //     aggregate — they never wait on another tile; a CHAIN warp walks the tiles in order, looks
        //     back over aggregates that are already out, publishes the inclusive prefix and hands the
        //     exclusive prefix to the scanners: the inter-tile chain advances at data-arrival time and
        //     never waits for the heavy work (nor the heavy work for the chain, unless it catches up);
        //   * SCANNER warps pull the stage into registers (LDS.128), scan it with shuffles, pick up the
        //     tile's exclusive prefix from the chain warp and store straight from registers.
        // Tiles are consumed by a CTA in ticket order, so forward progress holds as before.  A partial
        // tile bypasses the ring (guarded loads from global memory).
        template<typename T, int THREADS, int VPT, int STAGES>
        __global__ void __launch_bounds__(THREADS + 128)
            scan_b32_tma_kernel(T* __restrict__ data, size_t count, uint32_t tiles_per_part, uint32_t total_tiles,
                                uint32_t* ticket, uint64_t* state, const T* __restrict__ init)
        {
            constexpr int TILE = THREADS * VPT * 4;
            constexpr int WARPS = THREADS / 32; // scanner warps; then 1 producer, 2 aggregator, 1 chain warp
            constexpr int AGG_WARPS = 2;
            constexpr int WARP_ELEMS = VPT * 128;
            static_assert(WARPS <= 32, "one warp scans the warp totals");

            extern __shared__ __align__(128) unsigned char smem_raw[];
            T* ring = reinterpret_cast<T*>(smem_raw); // [STAGES][TILE]
            // The prefix hand-off uses 2*STAGES slots: the chain warp of tile it+STAGES may finish before the
            // scanners have picked up tile it's prefix (it only waits for them to have READ the stage).
            constexpr int PSLOTS = 2 * STAGES;
            __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], agg_bar[PSLOTS], chain_bar[PSLOTS];
            __shared__ uint32_t s_stage_tile[STAGES], s_stage_staged[STAGES];
            __shared__ T s_slot_agg[PSLOTS], s_slot_prefix[PSLOTS];
            __shared__ T s_warp_total[2][WARPS];

            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (threadIdx.x == 0)
            {
                for (int i = 0; i < STAGES; i++)
                {
                    mbarrier_init(&full_bar[i], 1);
                    mbarrier_init(&empty_bar[i], WARPS + 1);
                }
                for (int i = 0; i < PSLOTS; i++)
                {
                    mbarrier_init(&agg_bar[i], 1);
                    mbarrier_init(&chain_bar[i], 1);
                }
                mbarrier_init_fence();
            }
            __syncthreads();

            if (warp == WARPS)
            {
                // ---- producer: one lane
                if (lane == 0)
                {
                    const uint64_t policy = l2_policy_evict_first();
                    for (uint32_t it = 0;; it++)
                    {
                        const uint32_t stage = it % STAGES;
                        if (it >= STAGES)
                            mbarrier_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                        const uint32_t tile = atomicAdd(ticket, 1u);
                        s_stage_tile[stage] = tile;
                        if (tile >= total_tiles)
                        {
                            mbarrier_arrive(&full_bar[stage]);
                            break;
                        }
                        const uint32_t part = tile / tiles_per_part;
                        const uint32_t tp = tile - part * tiles_per_part;
                        const size_t in_part = size_t(tp) * TILE;
                        const size_t base = size_t(part) * count + in_part;
                        const bool staged =
                            count - in_part >= size_t(TILE) && (reinterpret_cast<uintptr_t>(data + base) & 15) == 0;
                        s_stage_staged[stage] = staged ? 1u : 0u;
                        if (staged)
                        {
                            mbarrier_arrive_expect_tx(&full_bar[stage], TILE * 4);
                            tma_load_1d(ring + size_t(stage) * TILE, data + base, TILE * 4, &full_bar[stage], policy);
                        }
                        else
                            mbarrier_arrive(&full_bar[stage]);
                    }
                }
                return;
            }

            if (warp > WARPS && warp <= WARPS + AGG_WARPS)
            {
                // ---- aggregator warps (alternating tiles): reduce the stage the moment it lands and publish
                // the tile aggregate.  They never wait on other tiles, so every tile in flight anywhere on
                // the GPU has its aggregate out as soon as its data is on chip.
                const unsigned me = warp - WARPS - 1;
                const T seed = init ? *init : T(0);
                for (uint32_t it = 0;; it++)
                {
                    const uint32_t stage = it % STAGES;
                    mbarrier_wait(&full_bar[stage], (it / STAGES) & 1);
                    const uint32_t tile = s_stage_tile[stage];
                    if (tile >= total_tiles)
                        break;
                    if (it % AGG_WARPS != me)
                        continue;
                    const uint32_t part = tile / tiles_per_part;
                    const uint32_t tp = tile - part * tiles_per_part;
                    T acc = T(0);
                    if (s_stage_staged[stage])
                    {
                        const uint4* src = reinterpret_cast<const uint4*>(ring + size_t(stage) * TILE) + lane;
                        T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
#pragma unroll 8
                        for (int k = 0; k < TILE / 128; k++)
                        {
                            const uint4 r = src[k * 32];
                            a0 += from_bits<T>(r.x);
                            a1 += from_bits<T>(r.y);
                            a2 += from_bits<T>(r.z);
                            a3 += from_bits<T>(r.w);
                        }
                        acc = (a0 + a1) + (a2 + a3);
                    }
                    else
                    {
                        const size_t in_part = size_t(tp) * TILE;
                        const size_t base = size_t(part) * count + in_part;
                        const uint32_t valid = uint32_t(count - in_part < size_t(TILE) ? count - in_part : size_t(TILE));
                        for (uint32_t idx = lane; idx < valid; idx += 32)
                            acc += data[base + idx];
                    }
                    const T aggregate = warp_sum(acc);
                    __syncwarp();
                    if (lane == 0)
                    {
                        mbarrier_arrive(&empty_bar[stage]); // done reading the stage
                        st_relaxed_u64(&state[tile], tp == 0 ? (k_flag_inclusive | to_bits<T>(seed + aggregate))
                                                             : (k_flag_aggregate | to_bits<T>(aggregate)));
                        s_slot_agg[it % PSLOTS] = aggregate;
                        mbarrier_arrive(&agg_bar[it % PSLOTS]);
                    }
                }
                return;
            }

            if (warp > WARPS + AGG_WARPS)
            {
                // ---- chain warp: in tile order, look back, publish the inclusive prefix, hand the exclusive
                // prefix to the scanners.  Its predecessors' aggregates are out long before it asks.
                const T seed = init ? *init : T(0);
                for (uint32_t it = 0;; it++)
                {
                    const uint32_t stage = it % STAGES;
                    mbarrier_wait(&full_bar[stage], (it / STAGES) & 1);
                    const uint32_t tile = s_stage_tile[stage];
                    if (tile >= total_tiles)
                        break;
                    const uint32_t part = tile / tiles_per_part;
                    const uint32_t tp = tile - part * tiles_per_part;
                    T exclusive = seed;
                    if (tp != 0)
                        exclusive = lookback_walk<T>(state, tile, tp, lane);
                    // The aggregator must be done with the tile before the scanners may overwrite it in
                    // global memory (an unstaged tile is reduced straight from global memory).
                    mbarrier_wait(&agg_bar[it % PSLOTS], (it / PSLOTS) & 1);
                    if (tp != 0 && lane == 0)
                        st_relaxed_u64(&state[tile], k_flag_inclusive | to_bits<T>(exclusive + s_slot_agg[it % PSLOTS]));
                    if (lane == 0)
                    {
                        s_slot_prefix[it % PSLOTS] = exclusive;
                        mbarrier_arrive(&chain_bar[it % PSLOTS]);
                    }
                }
                return;
            }

            // ---- scanner warps
            for (uint32_t it = 0;; it++)
            {
                const uint32_t stage = it % STAGES;
                const uint32_t parity = (it / STAGES) & 1;
                mbarrier_wait(&full_bar[stage], parity);
                const uint32_t tile = s_stage_tile[stage];
                if (tile >= total_tiles)
                    break;
                const bool staged = s_stage_staged[stage] != 0;
                const uint32_t part = tile / tiles_per_part;
                const uint32_t tp = tile - part * tiles_per_part;
                const size_t in_part = size_t(tp) * TILE;
                const size_t base = size_t(part) * count + in_part;
                const uint32_t valid = uint32_t(count - in_part < size_t(TILE) ? count - in_part : size_t(TILE));
                const uint32_t my_off = warp * WARP_ELEMS + lane * 4; // + j * 128

                T x[VPT][4];
                if (staged)
                {
                    const T* src = ring + size_t(stage) * TILE + my_off;
#pragma unroll
                    for (int j = 0; j < VPT; j++)
                    {
                        const uint4 r = *reinterpret_cast<const uint4*>(src + j * 128);
                        x[j][0] = from_bits<T>(r.x);
                        x[j][1] = from_bits<T>(r.y);
                        x[j][2] = from_bits<T>(r.z);
                        x[j][3] = from_bits<T>(r.w);
                    }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < VPT; j++)
#pragma unroll
                        for (int c = 0; c < 4; c++)
                        {
                            const uint32_t idx = my_off + j * 128 + c;
                            x[j][c] = idx < valid ? data[base + idx] : T(0);
                        }
                }
                __syncwarp();
                if (lane == 0)
                    mbarrier_arrive(&empty_bar[stage]); // this warp no longer needs the stage

                T sum4[VPT], inc[VPT], ex[VPT];
#pragma unroll
                for (int j = 0; j < VPT; j++)
                    inc[j] = sum4[j] = x[j][0] + x[j][1] + x[j][2] + x[j][3];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
#pragma unroll
                    for (int j = 0; j < VPT; j++)
                    {
                        T t = __shfl_up_sync(k_full_mask, inc[j], o);
                        if (lane >= unsigned(o))
                            inc[j] += t;
                    }
                T chunk_base = T(0);
#pragma unroll
                for (int j = 0; j < VPT; j++)
                {
                    T total = __shfl_sync(k_full_mask, inc[j], 31);
                    T lane_ex;
                    if (is_float_type<T>::value)
                    {
                        lane_ex = __shfl_up_sync(k_full_mask, inc[j], 1);
                        if (lane == 0)
                            lane_ex = T(0);
                    }
                    else
                        lane_ex = inc[j] - sum4[j];
                    ex[j] = chunk_base + lane_ex;
                    chunk_base += total;
                }
                T* totals = s_warp_total[it & 1]; // double-buffered: one barrier per tile is enough
                if (lane == 0)
                    totals[warp] = chunk_base;
                named_barrier_sync(1, THREADS);
                // every warp derives its own prefix from the warp totals (no second barrier)
                T wt = lane < warp ? totals[lane] : T(0);
                const T warp_prefix = warp_sum(wt);

                mbarrier_wait(&chain_bar[it % PSLOTS], (it / PSLOTS) & 1); // the chain warp has the tile's prefix
                const T prefix = s_slot_prefix[it % PSLOTS] + warp_prefix;
                if (staged)
                {
#pragma unroll
                    for (int j = 0; j < VPT; j++)
                    {
                        T b = prefix + ex[j];
                        uint4 r;
                        r.x = to_bits<T>(b);
                        b += x[j][0];
                        r.y = to_bits<T>(b);
                        b += x[j][1];
                        r.z = to_bits<T>(b);
                        b += x[j][2];
                        r.w = to_bits<T>(b);
                        st_stream_v4(data + base + my_off + j * 128, r);
                    }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < VPT; j++)
                    {
                        T b = prefix + ex[j];
#pragma unroll
                        for (int c = 0; c < 4; c++)
                        {
                            const uint32_t idx = my_off + j * 128 + c;
                            if (idx < valid)
                                data[base + idx] = b;
                            b += x[j][c];
                        }
                    }
                }
            }
        }

        // ------------------------------------------------------------------------ wide element types
        template<typename S, int NC> struct alignas(sizeof(S) * NC) Elem
        {
            S c[NC];
        };

        template<typename S, int NC> __device__ __forceinline__ Elem<S, NC> elem_zero()
        {
            Elem<S, NC> e;
#pragma unroll
            for (int k = 0; k < NC; k++)
                e.c[k] = S(0);
            return e;
        }
        template<typename S, int NC> __device__ __forceinline__ void elem_add(Elem<S, NC>& a, const Elem<S, NC>& b)
        {
#pragma unroll
            for (int k = 0; k < NC; k++)
                a.c[k] += b.c[k];
        }
        template<typename S, int NC>
        __device__ __forceinline__ Elem<S, NC> elem_shfl_up(const Elem<S, NC>& a, int o)
        {
            Elem<S, NC> r;
#pragma unroll
            for (int k = 0; k < NC; k++)
                r.c[k] = __shfl_up_sync(k_full_mask, a.c[k], o);
            return r;
        }
        template<typename S, int NC> __device__ __forceinline__ Elem<S, NC> elem_shfl(const Elem<S, NC>& a, int src)
        {
            Elem<S, NC> r;
#pragma unroll
            for (int k = 0; k < NC; k++)
                r.c[k] = __shfl_sync(k_full_mask, a.c[k], src);
            return r;
        }
        template<typename S, int NC> __device__ __forceinline__ Elem<S, NC> elem_warp_sum(Elem<S, NC> a)
        {
#pragma unroll
            for (int k = 0; k < NC; k++)
                a.c[k] = warp_sum<S>(a.c[k]);
            return a;
        }
        template<typename S, int NC>
        __device__ __forceinline__ Elem<S, NC> elem_warp_inclusive_scan(Elem<S, NC> v, unsigned lane)
        {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                Elem<S, NC> t = elem_shfl_up(v, o);
                if (lane >= unsigned(o))
                    elem_add(v, t);
            }
            return v;
        }
        template<typename S, int NC> __device__ __forceinline__ Elem<S, NC> elem_ldcg(const Elem<S, NC>* p)
        {
            Elem<S, NC> r;
            const S* q = reinterpret_cast<const S*>(p);
#pragma unroll
            for (int k = 0; k < NC; k++)
                r.c[k] = __ldcg(q + k);
            return r;
        }

        // state layout: flags u32[tiles] | aggregates E[tiles] | inclusives E[tiles]
        template<typename S, int NC, int THREADS, int IPT>
        __global__ void __launch_bounds__(THREADS)
            scan_wide_kernel(Elem<S, NC>* __restrict__ data, size_t count, uint32_t tiles_per_part, uint32_t* ticket,
                             uint32_t* flags, Elem<S, NC>* aggregates, Elem<S, NC>* inclusives,
                             const Elem<S, NC>* __restrict__ init)
        {
            using E = Elem<S, NC>;
            constexpr int TILE = THREADS * IPT;
            constexpr int WARPS = THREADS / 32;
            constexpr int WARP_ELEMS = IPT * 32;

            __shared__ E s_warp_total[WARPS];
            __shared__ E s_warp_prefix[WARPS];
            __shared__ E s_tile_prefix;
            __shared__ uint32_t s_tile;

            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (threadIdx.x == 0)
                s_tile = atomicAdd(ticket, 1u);
            __syncthreads();
            const uint32_t tile = s_tile;
            const uint32_t part = tile / tiles_per_part;
            const uint32_t tp = tile - part * tiles_per_part;
            const size_t in_part = size_t(tp) * TILE;
            const size_t base = size_t(part) * count + in_part;
            const uint32_t valid = uint32_t(count - in_part < size_t(TILE) ? count - in_part : size_t(TILE));
            const uint32_t my_off = warp * WARP_ELEMS + lane; // + j * 32

            E x[IPT], ex[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++)
            {
                uint32_t idx = my_off + j * 32;
                x[j] = idx < valid ? data[base + idx] : elem_zero<S, NC>();
            }
            E chunk_base = elem_zero<S, NC>();
#pragma unroll
            for (int j = 0; j < IPT; j++)
            {
                E inc = elem_warp_inclusive_scan(x[j], lane);
                E total = elem_shfl(inc, 31);
                E lane_ex = elem_shfl_up(inc, 1);
                if (lane == 0)
                    lane_ex = elem_zero<S, NC>();
                ex[j] = chunk_base;
                elem_add(ex[j], lane_ex);
                elem_add(chunk_base, total);
            }
            if (lane == 0)
                s_warp_total[warp] = chunk_base;
            __syncthreads();

            if (warp == 0)
            {
                E wt = lane < WARPS ? s_warp_total[lane] : elem_zero<S, NC>();
                E winc = elem_warp_inclusive_scan(wt, lane);
                E wex = elem_shfl_up(winc, 1);
                if (lane == 0)
                    wex = elem_zero<S, NC>();
                if (lane < WARPS)
                    s_warp_prefix[lane] = wex;
                const E aggregate = elem_shfl(winc, 31);

                E exclusive = (tp == 0 && init) ? *init : elem_zero<S, NC>();
                if (tp == 0)
                {
                    if (lane == 0)
                    {
                        E inclusive = exclusive;
                        elem_add(inclusive, aggregate);
                        inclusives[tile] = inclusive;
                        st_release_u32(&flags[tile], 2u);
                    }
                }
                else
                {
                    if (lane == 0)
                    {
                        aggregates[tile] = aggregate;
                        st_release_u32(&flags[tile], 1u);
                    }
                    uint32_t remaining = tp;
                    uint32_t pred = tile - 1;
                    while (true)
                    {
                        const bool in_range = lane < remaining;
                        uint32_t flag, inclusive_mask;
                        while (true)
                        {
                            flag = in_range ? ld_acquire_u32(&flags[pred - lane]) : 2u;
                            const uint32_t empty_mask = __ballot_sync(k_full_mask, flag == 0);
                            inclusive_mask = __ballot_sync(k_full_mask, flag == 2);
                            const uint32_t need =
                                inclusive_mask ? ((2u << (__ffs(inclusive_mask) - 1)) - 1u) : k_full_mask;
                            if ((empty_mask & need) == 0)
                                break;
                        }
                        const uint32_t first = inclusive_mask ? uint32_t(__ffs(inclusive_mask) - 1) : 31u;
                        E contrib = elem_zero<S, NC>();
                        if (in_range && lane <= first)
                            contrib = elem_ldcg(flag == 2 ? &inclusives[pred - lane] : &aggregates[pred - lane]);
                        elem_add(exclusive, elem_warp_sum(contrib));
                        if (inclusive_mask)
                            break;
                        pred -= 32;
                        remaining -= 32;
                    }
                    if (lane == 0)
                    {
                        E inclusive = exclusive;
                        elem_add(inclusive, aggregate);
                        inclusives[tile] = inclusive;
                        st_release_u32(&flags[tile], 2u);
                    }
                }
                if (lane == 0)
                    s_tile_prefix = exclusive;
            }
            __syncthreads();

            E prefix = s_tile_prefix;
            elem_add(prefix, s_warp_prefix[warp]);
#pragma unroll
            for (int j = 0; j < IPT; j++)
            {
                uint32_t idx = my_off + j * 32;
                E b = prefix;
                elem_add(b, ex[j]);
                if (idx < valid)
                    data[base + idx] = b;
            }
        }

        // ------------------------------------------------------------------------------------ host side
        struct ScanPlan
        {
            uint32_t tile;           // elements per tile
            uint32_t tiles_per_part; // ceil(count / tile)
            uint64_t total_tiles;
            int variant;             // index into the tile-shape table of the element class
        };

        struct ScanShape
        {
            int threads, per_thread; // per_thread: 16-byte vectors (4-byte types) or elements (wide types)
        };
        // 4-byte element types: {threads, vectors per thread}; tile = threads * vpt * 4 elements
        // ids 0..5: one tile per CTA;  ids 6..: persistent TMA-pipelined kernel (threads exclude the producer warp)
        constexpr ScanShape k_b32_shapes[] = {{256, 4}, {64, 2},  {512, 4}, {512, 8}, {1024, 4}, {256, 8},
                                              {256, 8}, {512, 8}, {256, 8}, {512, 4}, {256, 4},  {512, 8}};
        constexpr int k_b32_stages[] = {0, 0, 0, 0, 0, 0, 3, 3, 2, 3, 4, 2};
        constexpr int k_b32_default = 6, k_b32_simple = 3, k_b32_small = 1, k_num_b32_shapes = 12;
        constexpr size_t k_persistent_min_elems = size_t(1) << 22; // below this the simple kernel is as good
        constexpr int k_wide_threads = 256, k_wide_ipt = 4; // 1024-element tiles
        constexpr int k_wide_small_threads = 64, k_wide_small_ipt = 2;

        int scan_env_int(const char* name, int fallback)
        {
            const char* v = std::getenv(name);
            return v && *v ? std::atoi(v) : fallback;
        }

        // `stageable`: the buffer is 16-byte aligned and so is every partition start, i.e. whole tiles can be
        // TMA-staged.  Otherwise (and for small inputs) the one-tile-per-CTA kernel is used.
        ScanPlan make_plan(size_t count, size_t num_partitions, bool wide, bool stageable)
        {
            static const int forced = scan_env_int("GLU_SCAN_CONFIG", -1); // tuning sweeps only
            static const size_t persistent_min =
                size_t(scan_env_int("GLU_SCAN_PERSISTENT_MIN", int(k_persistent_min_elems)));
            ScanPlan p;
            uint32_t big, small;
            int big_variant = 0;
            if (wide)
            {
                big = k_wide_threads * k_wide_ipt;
                small = k_wide_small_threads * k_wide_small_ipt;
            }
            else
            {
                big_variant = (forced >= 0 && forced < k_num_b32_shapes && forced != k_b32_small) ? forced : k_b32_default;
                if (k_b32_stages[big_variant] > 0 && (!stageable || count * num_partitions < persistent_min))
                    big_variant = k_b32_simple;
                big = k_b32_shapes[big_variant].threads * k_b32_shapes[big_variant].per_thread * 4;
                small = k_b32_shapes[k_b32_small].threads * k_b32_shapes[k_b32_small].per_thread * 4;
            }
            // short segments: a big tile would be mostly padding
            const bool use_small = num_partitions > 1 && count <= big / 2;
            p.variant = use_small ? 1 : big_variant;
            p.tile = use_small ? small : big;
            p.tiles_per_part = uint32_t((count + p.tile - 1) / p.tile);
            p.total_tiles = uint64_t(p.tiles_per_part) * num_partitions;
            return p;
        }

        size_t state_bytes(const ScanPlan& p, size_t elem_size)
        {
            if (elem_size == 4)
                return align_up(p.total_tiles * sizeof(uint64_t), k_tmp_align);
            return align_up(p.total_tiles * sizeof(uint32_t), k_tmp_align) +
                   2 * align_up(p.total_tiles * elem_size, k_tmp_align);
        }

        template<typename T, int THREADS, int VPT, int MIN_BLOCKS>
        int launch_b32_shape(T* data, size_t count, const ScanPlan& p, uint32_t* ticket, uint64_t* state, cudaStream_t s,
                             const T* init)
        {
            static const bool use_ticket = scan_env_int("GLU_SCAN_TICKET", 1) != 0;
            static const int debug_no_lookback = scan_env_int("GLU_SCAN_DEBUG_NO_LOOKBACK", 0);
            ScopedKernelProfile prof(GLU_KERNEL_SCAN, s);
            if (use_ticket)
                scan_b32_kernel<T, THREADS, VPT, MIN_BLOCKS, true><<<unsigned(p.total_tiles), THREADS, 0, s>>>(
                    data, count, p.tiles_per_part, ticket, state, debug_no_lookback, init);
            else
                scan_b32_kernel<T, THREADS, VPT, MIN_BLOCKS, false><<<unsigned(p.total_tiles), THREADS, 0, s>>>(
                    data, count, p.tiles_per_part, ticket, state, debug_no_lookback, init);
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }

        template<typename T, int THREADS, int VPT, int STAGES>
        int launch_b32_tma(T* data, size_t count, const ScanPlan& p, uint32_t* ticket, uint64_t* state, cudaStream_t s,
                           const T* init)
        {
            auto kernel = scan_b32_tma_kernel<T, THREADS, VPT, STAGES>;
            constexpr size_t smem = size_t(STAGES) * THREADS * VPT * 16;
            static std::atomic<bool> configured[64]; // per device; set after the attribute call (idempotent, so a race only repeats it)
            static int ctas_per_sm[64] = {};
            int dev = 0;
            GLU_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 64)
                return GLU_ERROR_INVALID_ARGUMENT;
            if (!configured[dev].load(std::memory_order_acquire))
            {
                GLU_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                GLU_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[dev], kernel, THREADS + 128, smem));
                configured[dev].store(true, std::memory_order_release);
            }
            const uint64_t resident = uint64_t(current_sm_count()) * uint64_t(ctas_per_sm[dev] > 0 ? ctas_per_sm[dev] : 1);
            const unsigned grid = unsigned(p.total_tiles < resident ? p.total_tiles : resident);
            ScopedKernelProfile prof(GLU_KERNEL_SCAN, s);
            kernel<<<grid, THREADS + 128, smem, s>>>(data, count, p.tiles_per_part, uint32_t(p.total_tiles), ticket, state, init);
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }

        template<typename T>
        int launch_b32(void* d_data, size_t count, const ScanPlan& p, void* d_tmp, cudaStream_t s, const void* d_init)
        {
            const T* init = static_cast<const T*>(d_init);
            uint32_t* ticket = static_cast<uint32_t*>(d_tmp);
            uint64_t* state = reinterpret_cast<uint64_t*>(static_cast<char*>(d_tmp) + k_tmp_align);
            T* data = static_cast<T*>(d_data);
            GLU_CUDA_TRY(cudaMemsetAsync(d_tmp, 0, k_tmp_align + state_bytes(p, 4), s));
            switch (p.variant)
            {
            case 0: return launch_b32_shape<T, 256, 4, 1>(data, count, p, ticket, state, s, init);
            case 1: return launch_b32_shape<T, 64, 2, 1>(data, count, p, ticket, state, s, init);
            case 2: return launch_b32_shape<T, 512, 4, 1>(data, count, p, ticket, state, s, init);
            case 3: return launch_b32_shape<T, 512, 8, 2>(data, count, p, ticket, state, s, init);
            case 4: return launch_b32_shape<T, 1024, 4, 2>(data, count, p, ticket, state, s, init);
            case 5: return launch_b32_shape<T, 256, 8, 4>(data, count, p, ticket, state, s, init);
            case 6: return launch_b32_tma<T, 256, 8, 3>(data, count, p, ticket, state, s, init);
            case 7: return launch_b32_tma<T, 512, 8, 3>(data, count, p, ticket, state, s, init);
            case 8: return launch_b32_tma<T, 256, 8, 2>(data, count, p, ticket, state, s, init);
            case 9: return launch_b32_tma<T, 512, 4, 3>(data, count, p, ticket, state, s, init);
            case 10: return launch_b32_tma<T, 256, 4, 4>(data, count, p, ticket, state, s, init);
            default: return launch_b32_tma<T, 512, 8, 2>(data, count, p, ticket, state, s, init);
            }
        }

        template<typename S, int NC>
        int launch_wide(void* d_data, size_t count, const ScanPlan& p, void* d_tmp, cudaStream_t s, const void* d_init)
        {
            using E = Elem<S, NC>;
            char* base = static_cast<char*>(d_tmp);
            uint32_t* ticket = reinterpret_cast<uint32_t*>(base);
            uint32_t* flags = reinterpret_cast<uint32_t*>(base + k_tmp_align);
            size_t flag_bytes = align_up(p.total_tiles * sizeof(uint32_t), k_tmp_align);
            size_t val_bytes = align_up(p.total_tiles * sizeof(E), k_tmp_align);
            E* aggregates = reinterpret_cast<E*>(base + k_tmp_align + flag_bytes);
            E* inclusives = reinterpret_cast<E*>(base + k_tmp_align + flag_bytes + val_bytes);
            GLU_CUDA_TRY(cudaMemsetAsync(d_tmp, 0, k_tmp_align + flag_bytes, s));
            ScopedKernelProfile prof(GLU_KERNEL_SCAN, s);
            if (p.variant == 0)
                scan_wide_kernel<S, NC, k_wide_threads, k_wide_ipt><<<unsigned(p.total_tiles), k_wide_threads, 0, s>>>(
                    static_cast<E*>(d_data), count, p.tiles_per_part, ticket, flags, aggregates, inclusives,
                    static_cast<const E*>(d_init));
            else
                scan_wide_kernel<S, NC, k_wide_small_threads, k_wide_small_ipt>
                    <<<unsigned(p.total_tiles), k_wide_small_threads, 0, s>>>(
                        static_cast<E*>(d_data), count, p.tiles_per_part, ticket, flags, aggregates, inclusives,
                    static_cast<const E*>(d_init));
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }
    } // namespace
} // namespace glu_b200

using namespace glu_b200;

extern "C" size_t glu_scan_exclusive_tmp_bytes(size_t count, size_t num_partitions, int data_type)
{
    DataTypeInfo info;
    if (!data_type_info(data_type, &info) || count == 0 || num_partitions == 0)
        return 0;
    size_t esz = info.scalar_size * info.ncomp;
    const size_t a = state_bytes(make_plan(count, num_partitions, esz != 4, true), esz);
    const size_t b = state_bytes(make_plan(count, num_partitions, esz != 4, false), esz);
    return k_tmp_align + (a > b ? a : b);
}

extern "C" int glu_scan_exclusive(void* d_data, size_t count, size_t num_partitions, int data_type, void* d_tmp,
                                  size_t tmp_bytes, glu_stream_t stream)
{
    return glu_scan_exclusive_init(d_data, count, num_partitions, data_type, nullptr, d_tmp, tmp_bytes, stream);
}

extern "C" int glu_scan_exclusive_init(void* d_data, size_t count, size_t num_partitions, int data_type,
                                       const void* d_init, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    DataTypeInfo info;
    if (!data_type_info(data_type, &info))
        return GLU_ERROR_INVALID_DATA_TYPE;
    if (!d_data || count == 0 || num_partitions == 0) // glu/BlellochScan.hpp:132-135
        return GLU_ERROR_INVALID_ARGUMENT;
    const size_t esz = info.scalar_size * info.ncomp;
    if (reinterpret_cast<uintptr_t>(d_data) % esz != 0 || reinterpret_cast<uintptr_t>(d_init) % esz != 0)
        return GLU_ERROR_MISALIGNED;
    if (count > (size_t(1) << 40) / esz / num_partitions)
        return GLU_ERROR_COUNT_TOO_LARGE;
    const bool stageable =
        reinterpret_cast<uintptr_t>(d_data) % 16 == 0 && (num_partitions == 1 || (count * esz) % 16 == 0);
    ScanPlan p = make_plan(count, num_partitions, esz != 4, stageable);
    if (p.total_tiles >= (uint64_t(1) << 31))
        return GLU_ERROR_COUNT_TOO_LARGE;
    if (!d_tmp || tmp_bytes < k_tmp_align + state_bytes(p, esz))
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (data_type)
    {
    case GLU_DATA_TYPE_UINT:
    case GLU_DATA_TYPE_INT: return launch_b32<uint32_t>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_FLOAT: return launch_b32<float>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_DOUBLE: return launch_wide<double, 1>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_VEC2: return launch_wide<float, 2>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_VEC4: return launch_wide<float, 4>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_DVEC2: return launch_wide<double, 2>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_DVEC4: return launch_wide<double, 4>(d_data, count, p, d_tmp, s, d_init);
    case GLU_DATA_TYPE_UVEC2:
    case GLU_DATA_TYPE_IVEC2: return launch_wide<uint32_t, 2>(d_data, count, p, d_tmp, s, d_init);
    default: return launch_wide<uint32_t, 4>(d_data, count, p, d_tmp, s, d_init);
    }
}
