// glu_radix_sort_seg.cuh — SEGMENTED stable LSD sort: many independent sorts in one set of launches (included by
// glu_radix_sort.cu).  It is what "each GPU runs the local onesweep on the remaining 24 bits" (BASELINE.json north_star)
// needs after the MSD split of the multi-GPU sort: a rank receives up to 256 top-digit buckets, every bucket must be
// sorted by its low key bits, and buckets must not mix.  One histogram launch + ceil(bits / 8) digit passes for ALL
// segments: 4 + 16 * passes bytes of HBM traffic per pair (52 B for 24 bits) instead of the 68 B of a full 32-bit sort
// of the received range.
//
// Layout.  Segment s holds count[s] pairs (device-resident counts: the exchange plan writes them).  In the INPUT
// arrays segment s starts at a multiple of the tile size (glu_radix_sort_segment_tile(), 7680 pairs): element offset
// first_tile[s] * TILE with first_tile = exclusive scan of ceil(count / TILE) — so tile t of a digit pass is always
// elements [t * TILE, (t + 1) * TILE) (every tile is a bulk copy, tiles never straddle segments) and only the last tile
// of a segment is partial (its slots past `valid` are padding: never counted, never written).  Intermediate passes
// keep that layout; the LAST pass writes the compact layout (segment s at sum of the counts before it), so the result
// is one dense array, segment after segment.  The passes ping-pong between the caller's two array pairs A and B
// (input in A); an odd number of passes leaves the result in B.
//
// Kernels: seg_setup_kernel / seg_tile_info_kernel (counts -> first tiles, compact bases, per-tile {valid, segment,
// first tile}), seg_histogram_kernel (all digit places of all segments in one read of the keys, lane-private shared
// bins as in histogram_kernel), seg_offsets_kernel (counts -> per-(pass, segment) digit offsets incl. the output base),
// onesweep_kernel<..., SEG = true> (default; GLU_SEG_KERNEL=1: onesweep_ring_kernel<..., SEG = true>, glu_onesweep_ring.cuh):
// the chain CTAs keep ONE running prefix over all tiles; a tile subtracts the prefix row in front of its segment's
// first tile.
//
// Runs variant (glu_radix_sort_u32kv_segmented_runs, what the multi-GPU sort calls after its copy-engine all-to-all):
// the input of the FIRST pass is a sequence of tile-aligned runs placed anywhere in the A arrays — a segment is one or
// more consecutive runs (bucket b = its runs from source rank 0, 1, ...).  run_tile_info_kernel expands the caller's run
// table into the first pass's own tile descriptors and a tile map (tile t is read from tile map[t] of the input); the
// histogram and the first pass read through the map, everything after is the layout above.

namespace glu_b200
{
namespace
{
    // tile shapes of the segmented passes: the ring kernel's and the one-tile-per-CTA kernel's — the same 7680 pairs, so
    // the tile-aligned layout does not depend on which kernel runs (GLU_SEG_KERNEL: 1 = ring, 0 = one tile per CTA)
    constexpr int k_seg_threads = 480, k_seg_ipt = 16, k_seg_blocks = 2;
    constexpr int k_seg1_threads = 320, k_seg1_ipt = 24, k_seg1_blocks = 3;
    constexpr int k_seg_tile = k_seg_threads * k_seg_ipt;
    static_assert(k_seg_tile == k_seg1_threads * k_seg1_ipt, "one layout for both kernels");
    constexpr int k_max_segments = 256;
    constexpr int k_seg_hist_group = 2; // tiles whose loads are in flight together in seg_histogram_kernel

    struct SegLayout
    {
        size_t max_tiles;
        size_t off_seg;   // [4][k_max_segments] u32: first_tile, compact base, (spare), (spare)
        size_t off_info;  // [max_tiles] uint2 {valid | seg << 24, first tile of the segment}
        size_t off_info0; // [max_tiles] uint2: the same for the FIRST pass of the runs variant (input made of runs)
        size_t off_map0;  // [max_tiles] u32: tile of the input arrays that tile t of the first pass is read from
        size_t off_hist;  // [k_max_passes][k_max_segments][256] u32
        size_t off_lookback; // per pass: max_tiles count rows + max_tiles prefix rows
        size_t zero_begin, zero_bytes; // tickets + num_tiles + histograms + look-back words: zeroed per sort
        size_t total;
    };

    SegLayout make_seg_layout(size_t max_tiles)
    {
        SegLayout l;
        l.max_tiles = max_tiles;
        l.off_seg = k_tmp_align; // [0, 256): 4 tickets, num_tiles at word 8
        l.off_info = l.off_seg + 4 * k_max_segments * sizeof(uint32_t);
        l.off_info0 = align_up(l.off_info + max_tiles * sizeof(uint2), k_tmp_align);
        l.off_map0 = align_up(l.off_info0 + max_tiles * sizeof(uint2), k_tmp_align);
        l.off_hist = align_up(l.off_map0 + max_tiles * sizeof(uint32_t), k_tmp_align);
        l.off_lookback = l.off_hist + size_t(k_max_passes) * k_max_segments * k_radix * sizeof(uint32_t);
        l.total = align_up(l.off_lookback + 2 * size_t(k_max_passes) * max_tiles * k_radix * sizeof(uint32_t), k_tmp_align);
        l.zero_begin = 0;
        l.zero_bytes = l.total;
        return l;
    }

    // One CTA of 256 threads, thread s = segment s.  seg[0][s] = first tile, seg[1][s] = compact base; *num_tiles.
    __global__ void __launch_bounds__(k_max_segments, 1)
        seg_setup_kernel(const uint32_t* __restrict__ seg_count, uint32_t num_segments, uint32_t max_tiles, uint32_t* seg,
                         uint32_t* num_tiles)
    {
        __shared__ uint32_t s_tiles[k_max_segments / 32], s_elems[k_max_segments / 32];
        const unsigned s = threadIdx.x, lane = s & 31, warp = s >> 5;
        const uint32_t cnt = s < num_segments ? seg_count[s] : 0u;
        const uint32_t tiles = (cnt + uint32_t(k_seg_tile) - 1) / uint32_t(k_seg_tile);
        uint32_t it = tiles, ie = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t a = __shfl_up_sync(k_full_mask, it, o), b = __shfl_up_sync(k_full_mask, ie, o);
            if (lane >= unsigned(o))
            {
                it += a;
                ie += b;
            }
        }
        if (lane == 31)
        {
            s_tiles[warp] = it;
            s_elems[warp] = ie;
        }
        __syncthreads();
        uint32_t bt = 0, be = 0, total_tiles = 0;
        for (unsigned w = 0; w < k_max_segments / 32; w++)
        {
            if (w < warp)
            {
                bt += s_tiles[w];
                be += s_elems[w];
            }
            total_tiles += s_tiles[w];
        }
        seg[s] = bt + it - tiles;                 // first tile of segment s
        seg[k_max_segments + s] = be + ie - cnt;   // compact base of segment s
        if (s == 0)
            *num_tiles = total_tiles <= max_tiles ? total_tiles : 0u; // more tiles than the scratch holds: sort nothing
    }

    __global__ void __launch_bounds__(256)
        seg_tile_info_kernel(const uint32_t* __restrict__ seg_count, const uint32_t* __restrict__ seg, uint32_t num_segments,
                             const uint32_t* __restrict__ num_tiles, uint2* info)
    {
        const uint32_t t = blockIdx.x * 256u + threadIdx.x;
        if (t >= *num_tiles)
            return;
        // the last segment whose first tile is <= t and that is not empty: binary search over first_tile (non-decreasing)
        uint32_t lo = 0, hi = num_segments; // first_tile[lo] <= t < first_tile[hi] (hi == num_segments: +inf)
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) / 2;
            if (seg[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        // empty segments share their first tile with the next segment: `lo` is the last of them, i.e. the owner
        const uint32_t first = seg[lo];
        const uint32_t left = seg_count[lo] - (t - first) * uint32_t(k_seg_tile);
        const uint32_t valid = left < uint32_t(k_seg_tile) ? left : uint32_t(k_seg_tile);
        info[t] = make_uint2(valid | (lo << 24), first);
    }

    // Runs variant (glu_radix_sort_u32kv_segmented_runs): the input of the FIRST pass is a sequence of `num_runs` runs in
    // segment order (a segment = one or more consecutive runs), run r = count[r] pairs starting at tile phys[r] of the
    // input arrays — wherever the exchange put it.  runs: 5 rows of (num_runs + 1) words: first tile of the run in the
    // pass's own tile numbering (exclusive scan of ceil(count / TILE); entry num_runs = the total), phys, count, segment,
    // first tile (same numbering) of the run's segment.  Writes the first pass's tile descriptors and its tile map.
    __global__ void __launch_bounds__(256)
        run_tile_info_kernel(const uint32_t* __restrict__ runs, uint32_t num_runs, uint32_t max_tiles, uint2* info0,
                             uint32_t* map0, uint32_t* num_tiles0)
    {
        const uint32_t stride = num_runs + 1;
        const uint32_t total = runs[num_runs] <= max_tiles ? runs[num_runs] : 0u; // more than the scratch holds: sort nothing
        const uint32_t t = blockIdx.x * 256u + threadIdx.x;
        if (t == 0)
            *num_tiles0 = total;
        if (t >= total)
            return;
        uint32_t lo = 0, hi = num_runs; // first[lo] <= t < first[hi]
        while (hi - lo > 1)
        {
            const uint32_t mid = (lo + hi) / 2;
            if (runs[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        // empty runs share their first tile with the next run: `lo` is the last of them, i.e. the one that owns tile t
        const uint32_t k = t - runs[lo];
        const uint32_t left = runs[2 * stride + lo] - k * uint32_t(k_seg_tile);
        const uint32_t valid = left < uint32_t(k_seg_tile) ? left : uint32_t(k_seg_tile);
        info0[t] = make_uint2(valid | (runs[3 * stride + lo] << 24), runs[4 * stride + lo]);
        map0[t] = runs[stride + lo] + k;
    }

    // hist[pass][segment][256] += digit counts of the valid keys of the tiles; a CTA takes a contiguous range of tiles
    // and flushes its shared bins whenever the segment changes.  Same lane-private bin layout as histogram_kernel.
    __global__ void __launch_bounds__(k_hist_threads, 1)
        seg_histogram_kernel(const uint32_t* __restrict__ keys, const uint2* __restrict__ info,
                             const uint32_t* __restrict__ d_num_tiles, int num_passes, uint32_t pre_shift, uint32_t key_mask,
                             uint32_t* hist, const uint32_t* __restrict__ tile_map)
    {
        extern __shared__ __align__(16) uint32_t s_hist[]; // [num_passes][k_radix][k_hist_copies]
        const uint32_t num_tiles = *d_num_tiles;
        const uint32_t per = (num_tiles + gridDim.x - 1) / gridDim.x;
        const uint32_t t_begin = blockIdx.x * per;
        const uint32_t t_end = t_begin + per < num_tiles ? t_begin + per : num_tiles;
        if (t_begin >= t_end)
            return;
        const unsigned lane = threadIdx.x & 31;
        uint32_t* mine = s_hist + lane;
        for (int i = threadIdx.x; i < num_passes * k_radix * k_hist_copies / 4; i += k_hist_threads)
            reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        auto flush = [&](uint32_t seg) {
            __syncthreads();
            for (int i = threadIdx.x; i < num_passes * k_radix; i += k_hist_threads)
            {
                uint32_t c = 0;
#pragma unroll
                for (int j = 0; j < k_hist_copies; j++)
                {
                    uint32_t& w = s_hist[i * k_hist_copies + ((j + lane) & (k_hist_copies - 1))]; // rotated: conflict free
                    c += w;
                    w = 0;
                }
                if (c)
                    atomicAdd(&hist[(size_t(i / k_radix) * k_max_segments + seg) * k_radix + (i % k_radix)], c);
            }
            __syncthreads();
        };
        // k_seg_hist_group tiles per round: all their loads are issued before the first count (two 128-bit loads per
        // thread and tile would leave the memory system idle most of the time)
        constexpr int G = k_seg_hist_group;
        constexpr int J = (k_seg_tile / 4 + k_hist_threads - 1) / k_hist_threads; // 128-bit units per thread and tile
        uint32_t cur_seg = info[t_begin].x >> 24;
        // (fetching the next round's descriptors a round ahead was measured slower: 0.298 against 0.274 ms at 2^28 keys)
        for (uint32_t t0 = t_begin; t0 < t_end; t0 += G)
        {
            uint2 ti[G];
            uint4 k[G][J];
#pragma unroll
            for (int g = 0; g < G; g++)
            {
                const bool in = t0 + g < t_end;
                ti[g] = in ? info[t0 + g] : make_uint2(0u, 0u); // valid = 0: nothing to count
                const uint32_t valid = ti[g].x & 0xffffffu;
                const uint32_t src = (in && tile_map) ? tile_map[t0 + g] : t0 + g;
                const uint4* body = reinterpret_cast<const uint4*>(keys + size_t(src) * k_seg_tile);
#pragma unroll
                for (int j = 0; j < J; j++)
                {
                    const uint32_t u = threadIdx.x + j * k_hist_threads;
                    k[g][j] = u * 4 < valid ? ld_stream_v4(body + u) : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++)
            {
                const uint32_t valid = ti[g].x & 0xffffffu, seg = ti[g].x >> 24;
                if (valid == 0) // uniform (past the end)
                    continue;
                if (seg != cur_seg) // uniform
                {
                    flush(cur_seg);
                    cur_seg = seg;
                }
#pragma unroll
                for (int j = 0; j < J; j++)
                {
                    const uint32_t u = threadIdx.x + j * k_hist_threads;
                    if (u * 4 >= valid)
                        continue;
                    const uint32_t kk[4] = {(k[g][j].x >> pre_shift) & key_mask, (k[g][j].y >> pre_shift) & key_mask,
                                            (k[g][j].z >> pre_shift) & key_mask, (k[g][j].w >> pre_shift) & key_mask};
                    if (u * 4 + 4 <= valid)
                    {
                        // all four keys count (every unit of a full tile — all but a segment's last tile): three
                        // instructions per count as in histogram_kernel; the kernel is issue-bound (ncu: 85 % issue
                        // slots, 47 % DRAM with per-key bounds checks)
#pragma unroll
                        for (int p = 0; p < k_max_passes; p++)
                        {
                            if (p >= num_passes) // uniform
                                break;
#pragma unroll
                            for (int c = 0; c < 4; c++)
                            {
                                const uint32_t d = __byte_perm(kk[c], 0u, 0x4440u + p);
                                atomicAdd(mine + p * k_radix * k_hist_copies + d * k_hist_copies, 1u);
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int c = 0; c < 4; c++)
                    {
                        if (u * 4 + c >= valid)
                            break;
#pragma unroll
                        for (int p = 0; p < k_max_passes; p++)
                        {
                            if (p >= num_passes)
                                break;
                            const uint32_t d = __byte_perm(kk[c], 0u, 0x4440u + p);
                            atomicAdd(mine + p * k_radix * k_hist_copies + d * k_hist_copies, 1u);
                        }
                    }
                }
            }
        }
        flush(cur_seg);
    }

    // grid (num_segments, num_passes), 256 threads: counts -> exclusive digit offsets + the output base of the pass
    // (last pass: the compact base of the segment, otherwise its tile-aligned base)
    __global__ void __launch_bounds__(k_radix, 1)
        seg_offsets_kernel(uint32_t* hist, const uint32_t* __restrict__ seg, int num_passes)
    {
        __shared__ uint32_t s_scan[k_radix / 32];
        const unsigned d = threadIdx.x, lane = d & 31, warp = d >> 5;
        const uint32_t s = blockIdx.x, p = blockIdx.y;
        uint32_t* row = hist + (size_t(p) * k_max_segments + s) * k_radix;
        const uint32_t c = row[d];
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(k_full_mask, inc, o);
            if (lane >= unsigned(o))
                inc += t;
        }
        if (lane == 31)
            s_scan[warp] = inc;
        __syncthreads();
        uint32_t off = 0;
        for (unsigned w = 0; w < warp; w++)
            off += s_scan[w];
        const uint32_t base = int(p) == num_passes - 1 ? seg[k_max_segments + s] : seg[s] * uint32_t(k_seg_tile);
        row[d] = base + off + inc - c;
    }
} // namespace
} // namespace glu_b200

extern "C" size_t glu_radix_sort_segment_tile(void) { return size_t(k_seg_tile); }

extern "C" size_t glu_radix_sort_u32kv_segmented_tmp_bytes(size_t max_tiles)
{
    if (max_tiles == 0 || max_tiles * size_t(k_seg_tile) > k_max_count)
        return 0;
    return make_seg_layout(max_tiles).total;
}

namespace
{
int seg_sort_impl(uint32_t* d_keys_a, uint32_t* d_vals_a, uint32_t* d_keys_b, uint32_t* d_vals_b, const uint32_t* d_seg_count,
                  size_t num_segments, size_t max_tiles, unsigned begin_bit, unsigned end_bit, const uint32_t* d_runs,
                  size_t num_runs, void* d_tmp, size_t tmp_bytes, glu_stream_t stream, int* result_in_b)
{
    if (!d_keys_a || !d_vals_a || !d_keys_b || !d_vals_b || !d_seg_count || num_segments == 0 ||
        num_segments > size_t(k_max_segments) || begin_bit >= end_bit || end_bit > 32 || max_tiles == 0)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (max_tiles * size_t(k_seg_tile) > k_max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    // every tile is a bulk copy: the arrays must be 16-byte aligned (cudaMalloc'ed arrays are)
    if ((reinterpret_cast<uintptr_t>(d_keys_a) | reinterpret_cast<uintptr_t>(d_vals_a) | reinterpret_cast<uintptr_t>(d_keys_b) |
         reinterpret_cast<uintptr_t>(d_vals_b)) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(d_seg_count) % sizeof(uint32_t) != 0)
        return GLU_ERROR_MISALIGNED;
    const SegLayout l = make_seg_layout(max_tiles);
    if (!d_tmp || tmp_bytes < l.total)
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    const int sms = current_sm_count();
    if (sms <= 0)
        return GLU_ERROR_CUDA;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* tmp = static_cast<char*>(d_tmp);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(tmp);
    uint32_t* num_tiles = tickets + 8;
    uint32_t* seg = reinterpret_cast<uint32_t*>(tmp + l.off_seg);
    uint2* info = reinterpret_cast<uint2*>(tmp + l.off_info);
    // runs variant: the first pass (and the histogram) read the input through its own tile descriptors and tile map
    uint32_t* num_tiles0 = d_runs ? tickets + 9 : num_tiles;
    uint2* info0 = d_runs ? reinterpret_cast<uint2*>(tmp + l.off_info0) : info;
    uint32_t* map0 = d_runs ? reinterpret_cast<uint32_t*>(tmp + l.off_map0) : nullptr;
    uint32_t* hist = reinterpret_cast<uint32_t*>(tmp + l.off_hist);
    uint32_t* lookback = reinterpret_cast<uint32_t*>(tmp + l.off_lookback);
    const PassPlan plan = make_bit_plan(begin_bit, int(end_bit - begin_bit));
    const uint32_t S = uint32_t(num_segments);

    // tickets, histograms and the look-back words of the passes that run
    GLU_CUDA_TRY(cudaMemsetAsync(tmp, 0, k_tmp_align, s));
    GLU_CUDA_TRY(cudaMemsetAsync(hist, 0,
                                 l.off_lookback - l.off_hist + 2 * size_t(plan.num_passes) * max_tiles * k_radix * sizeof(uint32_t),
                                 s));
    seg_setup_kernel<<<1, k_max_segments, 0, s>>>(d_seg_count, S, uint32_t(max_tiles), seg, num_tiles);
    GLU_LAUNCH_CHECK();
    seg_tile_info_kernel<<<unsigned((max_tiles + 255) / 256), 256, 0, s>>>(d_seg_count, seg, S, num_tiles, info);
    GLU_LAUNCH_CHECK();
    if (d_runs)
    {
        run_tile_info_kernel<<<unsigned((max_tiles + 255) / 256), 256, 0, s>>>(d_runs, uint32_t(num_runs), uint32_t(max_tiles),
                                                                                info0, map0, num_tiles0);
        GLU_LAUNCH_CHECK();
    }
    {
        const size_t smem = size_t(plan.num_passes) * k_radix * k_hist_copies * sizeof(uint32_t);
        static std::atomic<bool> configured[64];
        int dev = 0;
        GLU_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 64)
            return GLU_ERROR_INVALID_ARGUMENT;
        if (!configured[dev].load(std::memory_order_acquire))
        {
            GLU_CUDA_TRY(cudaFuncSetAttribute(seg_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              int(k_max_passes * k_radix * k_hist_copies * sizeof(uint32_t))));
            configured[dev].store(true, std::memory_order_release);
        }
        const unsigned grid = unsigned(max_tiles < size_t(sms) ? max_tiles : size_t(sms));
        ScopedKernelProfile prof(GLU_KERNEL_SORT_HISTOGRAM, s);
        seg_histogram_kernel<<<grid, k_hist_threads, smem, s>>>(d_keys_a, info0, num_tiles0, plan.num_passes, plan.begin_bit,
                                                                plan.key_mask, hist, map0);
        GLU_LAUNCH_CHECK();
    }
    seg_offsets_kernel<<<dim3(S, unsigned(plan.num_passes)), k_radix, 0, s>>>(hist, seg, plan.num_passes);
    GLU_LAUNCH_CHECK();

    uint32_t* kbuf[2] = {d_keys_a, d_keys_b};
    uint32_t* vbuf[2] = {d_vals_a, d_vals_b};
    for (int p = 0; p < plan.num_passes; p++)
    {
        uint32_t* lb = lookback + 2 * size_t(p) * max_tiles * k_radix;
        static const int use_ring_env = env_int("GLU_SEG_KERNEL", 0);
        const bool first_of_runs = d_runs && p == 0;
        const bool use_ring = use_ring_env && !d_runs; // the tile map exists in the one-tile-per-CTA kernel only
        const uint32_t n_padded = uint32_t(max_tiles * size_t(k_seg_tile));
        const uint32_t* offsets = hist + size_t(p) * k_max_segments * k_radix;
        const int rc =
            use_ring ? launch_ring<k_seg_threads, k_seg_ipt, k_seg_blocks, Rank_Ballot, 0, 0, true>(
                           kbuf[p & 1], vbuf[p & 1], kbuf[(p + 1) & 1], vbuf[(p + 1) & 1], n_padded, plan.shift[p],
                           plan.mask[p], offsets, lb, tickets + p, unsigned(max_tiles), s, num_tiles, info)
                     : launch_sweep<k_seg1_threads, k_seg1_ipt, k_seg1_blocks, Rank_Ballot, false, false, 0, true>(
                           kbuf[p & 1], vbuf[p & 1], kbuf[(p + 1) & 1], vbuf[(p + 1) & 1], n_padded, plan.shift[p],
                           plan.mask[p], offsets, lb, tickets + p, unsigned(max_tiles), s,
                           first_of_runs ? num_tiles0 : num_tiles, nullptr, nullptr, nullptr, first_of_runs ? info0 : info,
                           first_of_runs ? map0 : nullptr);
        if (rc != GLU_SUCCESS)
            return rc;
    }
    if (result_in_b)
        *result_in_b = plan.num_passes & 1;
    return GLU_SUCCESS;
}
} // namespace

extern "C" int glu_radix_sort_u32kv_segmented(uint32_t* d_keys_a, uint32_t* d_vals_a, uint32_t* d_keys_b, uint32_t* d_vals_b,
                                              const uint32_t* d_seg_count, size_t num_segments, size_t max_tiles,
                                              unsigned begin_bit, unsigned end_bit, void* d_tmp, size_t tmp_bytes,
                                              glu_stream_t stream, int* result_in_b)
{
    return seg_sort_impl(d_keys_a, d_vals_a, d_keys_b, d_vals_b, d_seg_count, num_segments, max_tiles, begin_bit, end_bit,
                         nullptr, 0, d_tmp, tmp_bytes, stream, result_in_b);
}

extern "C" int glu_radix_sort_u32kv_segmented_runs(uint32_t* d_keys_a, uint32_t* d_vals_a, uint32_t* d_keys_b,
                                                   uint32_t* d_vals_b, const uint32_t* d_seg_count, size_t num_segments,
                                                   size_t max_tiles, unsigned begin_bit, unsigned end_bit,
                                                   const uint32_t* d_runs, size_t num_runs, void* d_tmp, size_t tmp_bytes,
                                                   glu_stream_t stream, int* result_in_b)
{
    if (!d_runs || num_runs == 0 || num_runs > size_t(1) << 16 || reinterpret_cast<uintptr_t>(d_runs) % sizeof(uint32_t) != 0)
        return GLU_ERROR_INVALID_ARGUMENT;
    return seg_sort_impl(d_keys_a, d_vals_a, d_keys_b, d_vals_b, d_seg_count, num_segments, max_tiles, begin_bit, end_bit,
                         d_runs, num_runs, d_tmp, tmp_bytes, stream, result_in_b);
}
