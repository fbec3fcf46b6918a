// glu_oracle.cpp — CPU ORACLE for the glu hot path (Reduce / BlellochScan / RadixSort).
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library, and only as the checker
// or as the timed CPU baseline.  The product (libglu_b200.so) never links, loads or calls it.
//
// Parity status: PINNED.  The reference implementation itself (GLSL 4.60 compute shaders behind an
// OpenGL 4.6 context) cannot be built or run in this environment (no X11/GL/EGL — see DESIGN.md),
// so oracle/_ref does not exist.  Instead this file holds
//   (A) the reference's own definition of "correct": the std:: algorithms its test-suite compares
//       against (test/reduce_tests.cpp:155,174; test/blelloch_scan_tests.cpp:44,75;
//       test/radix_sort_tests.cpp:20-51), strengthened to std::stable_sort on (key,val) pairs;
//   (B) a dispatch-by-dispatch CPU restatement of the reference's host loops + shaders
//       (glu/Reduce.hpp:11-38,111-135; glu/BlellochScan.hpp:13-76,142-190;
//       glu/RadixSort.hpp:11-58,60-183,273-334) used to prove that (A) and the reference's real
//       algorithm agree, including on properties the reference's tests never check (values,
//       stability, bit 31, num_steps);
//   (C) the reference's input generator glu::Random (test/util/Random.hpp:12-38).
// tests/test_oracle.py checks (A), (B) and (C) against every known-answer constant in the
// reference's tests (4951 / 319200 / 1 / 99 / 505 / ...) and against golden vectors in
// tests/golden/ that were produced from the reference's generator.
//
// Build: make -C oracle   (g++ -O2 -fopenmp -shared -fPIC)  -> oracle/libglu_oracle.so

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#include <parallel/algorithm>
#endif

namespace
{
    // ---- reference int helpers (glu/gl_utils.hpp:279-302) -------------------------------------------------

    // glu/gl_utils.hpp:279-283 — div_ceil goes through double; exact for the sizes used here.
    inline size_t div_ceil(size_t n, size_t d) { return (size_t) std::ceil(double(n) / double(d)); }

    // glu/gl_utils.hpp:291-302 — 32-bit smear (0 -> 0).
    inline size_t next_power_of_2(size_t n)
    {
        n--;
        n |= n >> 1;
        n |= n >> 2;
        n |= n >> 4;
        n |= n >> 8;
        n |= n >> 16;
        n++;
        return n;
    }

    constexpr size_t k_num_threads = 1024; // m_num_threads in all three reference classes
    constexpr size_t k_subgroup = 32;      // hard-coded by `1 << (5 * depth)` glu/Reduce.hpp:26,123

    enum Op
    {
        Op_Sum = 0,
        Op_Mul,
        Op_Min,
        Op_Max
    }; // glu/Reduce.hpp:42-48

    template<typename T> inline T apply(int op, T a, T b)
    {
        switch (op)
        {
        case Op_Sum: return T(a + b);
        case Op_Mul: return T(a * b);
        case Op_Min: return b < a ? b : a;
        default: return a < b ? b : a;
        }
    }

    // ---- (B) Reduce: one dispatch per radix-32 tree level ---------------------------------------------------
    // glu/Reduce.hpp:121-134 (host loop) + :24-37 (shader).  `ncomp` interleaved components model the
    // vecN types (subgroup ops are component-wise).  Subgroup operation order is implementation
    // defined in GLSL; lane order is used here (exact for integers, tolerance for floats).
    template<typename T> void reduce_tree(T* data, size_t count, int ncomp, int op)
    {
        for (int depth = 0;; depth++)
        {
            size_t step = size_t(1) << (5 * depth);
            if (step >= count)
                break;
            size_t level_count = count >> (5 * depth);
            size_t num_workgroups = div_ceil(level_count, k_num_threads);
            for (size_t wg = 0; wg < num_workgroups; wg++)
                for (size_t sg = 0; sg < k_num_threads / k_subgroup; sg++)
                {
                    size_t subgroup_i = wg * k_num_threads + sg * k_subgroup;
                    size_t i0 = subgroup_i * step;
                    if (i0 >= count)
                        continue; // lane 0 inactive => nobody writes
                    for (int c = 0; c < ncomp; c++)
                    {
                        T r = data[i0 * ncomp + c];
                        for (size_t lane = 1; lane < k_subgroup; lane++)
                        {
                            size_t i = (subgroup_i + lane) * step;
                            if (i < count)
                                r = apply<T>(op, r, data[i * ncomp + c]);
                        }
                        data[i0 * ncomp + c] = r;
                    }
                }
        }
    }

    // ---- (B) BlellochScan: one dispatch per tree level --------------------------------------------------------
    // Host loops glu/BlellochScan.hpp:142-166 (upsweep) and :168-190 (downsweep); shaders :26-45, :59-75.
    template<typename T> void blelloch_scan(T* data, size_t count, size_t num_partitions)
    {
        // upsweep (+ "clear last" at every level)
        {
            int step = 1;
            int level_count = (int) count;
            while (true)
            {
                size_t num_workgroups = div_ceil((size_t) level_count, k_num_threads);
                for (size_t part = 0; part < num_partitions; part++)
                {
                    size_t end_i = (part + 1) * count;
                    for (size_t sg0 = 0; sg0 < num_workgroups * k_num_threads; sg0 += k_subgroup)
                    {
                        // a subgroup executes in lock-step: all reads (own + shuffled) precede the writes
                        T own[k_subgroup];
                        bool active[k_subgroup];
                        for (size_t lane = 0; lane < k_subgroup; lane++)
                        {
                            size_t i = part * count + (sg0 + lane) * size_t(step) + size_t(step) - 1;
                            active[lane] = i < end_i;
                            own[lane] = active[lane] ? data[i] : T(0);
                        }
                        for (size_t lane = 0; lane < k_subgroup; lane++)
                        {
                            if (!active[lane])
                                continue;
                            size_t i = part * count + (sg0 + lane) * size_t(step) + size_t(step) - 1;
                            T lval = lane > 0 ? own[lane - 1] : T(0); // subgroupShuffleUp(data[i], 1)
                            T r = T(own[lane] + lval);
                            if (i == end_i - 1)
                                data[i] = T(0); // clear last
                            else if (lane % 2 == 1)
                                data[i] = r;
                        }
                    }
                }
                step <<= 1;
                level_count >>= 1;
                if (level_count <= 1)
                    break;
            }
        }
        // downsweep
        {
            int step = (int) (next_power_of_2((size_t) (int) count) >> 1);
            size_t level_count = 1;
            while (true)
            {
                size_t num_workgroups = div_ceil(level_count, k_num_threads);
                for (size_t part = 0; part < num_partitions; part++)
                {
                    size_t end_i = (part + 1) * count;
                    for (size_t t = 0; t < num_workgroups * k_num_threads; t++)
                    {
                        size_t i = part * count + t * (size_t(step) << 1) + (size_t(step) - 1);
                        size_t next_i = i + size_t(step);
                        if (next_i < end_i)
                        {
                            T tmp = data[i];
                            data[i] = data[next_i];
                            data[next_i] = T(data[next_i] + tmp);
                        }
                        else if (i < end_i)
                            data[i] = T(0);
                    }
                }
                step >>= 1;
                level_count <<= 1;
                if (step == 0)
                    break;
            }
        }
    }

    // ---- (B) RadixSort: 8 x (count, 16-partition scan, reorder) ------------------------------------------------
    // glu/RadixSort.hpp:273-334 (host), :33-57 (count shader), :142-182 (reorder shader; the shared-memory
    // Blelloch prefix_sum :102-140 is an exclusive scan of the 1024 match flags, restated as a running count).
    // Returns the index (0 = caller buffers, 1 = internal scratch) of the buffer pair holding the result —
    // the reference leaves an odd-num_steps result in its scratch.
    int radix_sort_glsl(std::vector<uint32_t> (&kb)[2], std::vector<uint32_t> (&vb)[2], size_t count,
                        size_t num_steps)
    {
        if (count <= 1)
            return 0;
        size_t num_blocks = div_ceil(count, size_t(1024));
        size_t nbp2 = next_power_of_2(num_blocks);
        std::vector<uint32_t> block_count(next_power_of_2(16 * nbp2));
        uint32_t global_count[16];
        kb[1].assign(next_power_of_2(count), 0);
        vb[1].assign(next_power_of_2(count), 0);

        int step = 0;
        for (; step < 8;)
        {
            const std::vector<uint32_t>& sk = kb[step % 2];
            const std::vector<uint32_t>& sv = vb[step % 2];
            std::vector<uint32_t>& dk = kb[(step + 1) % 2];
            std::vector<uint32_t>& dv = vb[(step + 1) % 2];
            uint32_t shift = uint32_t(step) << 2;

            std::fill(block_count.begin(), block_count.end(), 0u);
            std::fill(global_count, global_count + 16, 0u);
            // counting dispatch
            for (size_t wg = 0; wg < num_blocks; wg++)
            {
                for (size_t t = 0; t < 1024; t++)
                {
                    size_t i = wg * 1024 + t;
                    if (i < count)
                        block_count[((sk[i] >> shift) & 0xf) * nbp2 + wg]++;
                }
                for (size_t r = 0; r < 16; r++)
                    global_count[r] += block_count[r * nbp2 + wg];
            }
            // m_blelloch_scan(block_count, nbp2, 16)
            blelloch_scan<uint32_t>(block_count.data(), nbp2, 16);
            // reorder dispatch
            uint32_t global_off[16];
            {
                uint32_t acc = 0; // subgroupExclusiveAdd over 16 lanes (:148-152)
                for (int r = 0; r < 16; r++)
                {
                    global_off[r] = acc;
                    acc += global_count[r];
                }
            }
            for (size_t wg = 0; wg < num_blocks; wg++)
                for (uint32_t radix = 0; radix < 16; radix++)
                {
                    uint32_t local = 0;
                    for (size_t t = 0; t < 1024; t++)
                    {
                        size_t i = wg * 1024 + t;
                        if (i < count && ((sk[i] >> shift) & 0xf) == radix)
                        {
                            size_t di = size_t(global_off[radix]) + block_count[radix * nbp2 + wg] + local;
                            dk[di] = sk[i];
                            dv[di] = sv[i];
                            local++;
                        }
                    }
                }
            ++step;
            if (size_t(step) == num_steps || step == 8)
                break;
        }
        return step % 2;
    }

    struct KeyLess
    {
        uint32_t mask;
        bool operator()(const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) const
        {
            return (a.first & mask) < (b.first & mask);
        }
    };

    inline uint32_t steps_mask(size_t num_steps)
    {
        // glu/RadixSort.hpp:289-333 — num_steps==0 or >=8 sorts all 32 bits; otherwise the low 4*num_steps bits.
        if (num_steps == 0 || num_steps >= 8)
            return 0xffffffffu;
        return (uint32_t(1) << (4 * num_steps)) - 1u;
    }
} // namespace

extern "C"
{
    int glu_oracle_version() { return 1; }

    int glu_oracle_max_threads()
    {
#if defined(_OPENMP)
        return omp_get_max_threads();
#else
        return 1;
#endif
    }

    // ---- (C) glu::Random — test/util/Random.hpp:12-38 ----------------------------------------------------------
    // seed==0 -> default-constructed std::minstd_rand; sample_int = engine() % (max-min) + min (half-open).
    // IntegerT = GLuint: the modulo is evaluated in uint_fast32_t (64-bit), then truncated.
    void glu_oracle_random_u32(uint64_t seed, size_t n, uint32_t min, uint32_t max, uint32_t* out)
    {
        std::minstd_rand engine = seed != 0 ? std::minstd_rand(seed) : std::minstd_rand();
        for (size_t i = 0; i < n; i++)
            out[i] = uint32_t((engine() % (max - min)) + min);
    }

    // Non-reference generators used by SURVEY.md §8(d) configs (true 32-bit keys, skew).
    void glu_oracle_mt19937_u32(uint32_t seed, size_t n, uint32_t* out)
    {
        std::mt19937 engine(seed);
        for (size_t i = 0; i < n; i++)
            out[i] = uint32_t(engine());
    }

    // ---- (A) std:: oracles -------------------------------------------------------------------------------------

    // test/reduce_tests.cpp:155,174 — std::accumulate(begin, end, GLuint(0)); generalised to the 4 operators.
    uint32_t glu_oracle_reduce_u32(const uint32_t* data, size_t n, int op)
    {
        switch (op)
        {
        case Op_Sum: return std::accumulate(data, data + n, uint32_t(0));
        case Op_Mul:
            return std::accumulate(data, data + n, uint32_t(1), [](uint32_t a, uint32_t b) { return uint32_t(a * b); });
        case Op_Min: return *std::min_element(data, data + n);
        default: return *std::max_element(data, data + n);
        }
    }

    int32_t glu_oracle_reduce_i32(const int32_t* data, size_t n, int op)
    {
        switch (op)
        {
        case Op_Sum:
            return (int32_t) std::accumulate(data, data + n, uint32_t(0),
                                             [](uint32_t a, int32_t b) { return uint32_t(a + uint32_t(b)); });
        case Op_Mul:
            return (int32_t) std::accumulate(data, data + n, uint32_t(1),
                                             [](uint32_t a, int32_t b) { return uint32_t(a * uint32_t(b)); });
        case Op_Min: return *std::min_element(data, data + n);
        default: return *std::max_element(data, data + n);
        }
    }

    // Floating reductions: accumulate in double (float input) / long double (double input); callers
    // compare with a stated tolerance (the reference uses +-0.1 absolute, test/reduce_tests.cpp:72).
    double glu_oracle_reduce_f32(const float* data, size_t n, int op)
    {
        switch (op)
        {
        case Op_Sum: return std::accumulate(data, data + n, 0.0, [](double a, float b) { return a + double(b); });
        case Op_Mul: return std::accumulate(data, data + n, 1.0, [](double a, float b) { return a * double(b); });
        case Op_Min: return double(*std::min_element(data, data + n));
        default: return double(*std::max_element(data, data + n));
        }
    }

    double glu_oracle_reduce_f64(const double* data, size_t n, int op)
    {
        switch (op)
        {
        case Op_Sum:
            return (double) std::accumulate(data, data + n, (long double) 0,
                                            [](long double a, double b) { return a + (long double) b; });
        case Op_Mul:
            return (double) std::accumulate(data, data + n, (long double) 1,
                                            [](long double a, double b) { return a * (long double) b; });
        case Op_Min: return *std::min_element(data, data + n);
        default: return *std::max_element(data, data + n);
        }
    }

    // sum of |x| in double: scale for the relative tolerance of float sums
    double glu_oracle_sum_abs_f32(const float* data, size_t n)
    {
        return std::accumulate(data, data + n, 0.0, [](double a, float b) { return a + std::fabs(double(b)); });
    }

    // test/blelloch_scan_tests.cpp:44,75 — std::exclusive_scan(begin, end, out, 0) per partition.
    void glu_oracle_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t count, size_t num_partitions)
    {
        for (size_t p = 0; p < num_partitions; p++)
            std::exclusive_scan(in + p * count, in + (p + 1) * count, out + p * count, uint32_t(0));
    }

    void glu_oracle_exclusive_scan_f32(const float* in, float* out, size_t count, size_t num_partitions)
    {
        // sequential float adds in double, rounded at the end: the comparison is tolerance-based
        for (size_t p = 0; p < num_partitions; p++)
        {
            double acc = 0;
            for (size_t i = 0; i < count; i++)
            {
                out[p * count + i] = float(acc);
                acc += double(in[p * count + i]);
            }
        }
    }

    void glu_oracle_exclusive_scan_f64(const double* in, double* out, size_t count, size_t num_partitions)
    {
        for (size_t p = 0; p < num_partitions; p++)
        {
            long double acc = 0;
            for (size_t i = 0; i < count; i++)
            {
                out[p * count + i] = double(acc);
                acc += (long double) in[p * count + i];
            }
        }
    }

    // test/radix_sort_tests.cpp:20-51 (is_sorted + permutation), strengthened per north_star to
    // std::stable_sort of (key,val) pairs comparing keys only.  num_steps follows glu/RadixSort.hpp:331.
    // threads<=1: std::stable_sort; threads>1: __gnu_parallel::stable_sort (same result, it is stable).
    void glu_oracle_stable_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, size_t num_steps, int threads)
    {
        std::vector<std::pair<uint32_t, uint32_t>> pairs(n);
        for (size_t i = 0; i < n; i++)
            pairs[i] = {keys[i], vals[i]};
        KeyLess less{steps_mask(num_steps)};
#if defined(_OPENMP)
        if (threads > 1)
        {
            omp_set_num_threads(threads);
            __gnu_parallel::stable_sort(pairs.begin(), pairs.end(), less);
        }
        else
#endif
            std::stable_sort(pairs.begin(), pairs.end(), less);
        for (size_t i = 0; i < n; i++)
        {
            keys[i] = pairs[i].first;
            vals[i] = pairs[i].second;
        }
    }

    // Oracle of glu_radix_sort_u32_ex (beyond the reference; PARITY UNPINNED by the reference,
    // which has neither key-only, bit-range nor descending sorts): std::stable_sort of the pairs comparing only key
    // bits [begin_bit, end_bit), with std::less (ascending) or std::greater (descending) on those bits.  vals may be
    // null (keys only).
    void glu_oracle_stable_sort_ex(uint32_t* keys, uint32_t* vals, size_t n, unsigned begin_bit, unsigned end_bit,
                                   int descending)
    {
        const unsigned bits = end_bit - begin_bit;
        const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
        std::vector<std::pair<uint32_t, uint32_t>> pairs(n);
        for (size_t i = 0; i < n; i++)
            pairs[i] = {keys[i], vals ? vals[i] : 0u};
        auto field = [=](uint32_t k) { return begin_bit >= 32 ? 0u : ((k >> begin_bit) & mask); };
        if (descending)
            std::stable_sort(pairs.begin(), pairs.end(),
                             [=](const auto& a, const auto& b) { return field(a.first) > field(b.first); });
        else
            std::stable_sort(pairs.begin(), pairs.end(),
                             [=](const auto& a, const auto& b) { return field(a.first) < field(b.first); });
        for (size_t i = 0; i < n; i++)
        {
            keys[i] = pairs[i].first;
            if (vals)
                vals[i] = pairs[i].second;
        }
    }

    // Same contract as above with the pair array already built (what bench.py times: the sort only).
    // Returns seconds spent inside stable_sort.
    double glu_oracle_time_stable_sort_pairs(const uint32_t* keys, const uint32_t* vals, size_t n, int threads)
    {
        std::vector<std::pair<uint32_t, uint32_t>> pairs(n);
        for (size_t i = 0; i < n; i++)
            pairs[i] = {keys[i], vals[i]};
        KeyLess less{0xffffffffu};
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
#if defined(_OPENMP)
        if (threads > 1)
        {
            omp_set_num_threads(threads);
            __gnu_parallel::stable_sort(pairs.begin(), pairs.end(), less);
        }
        else
#endif
            std::stable_sort(pairs.begin(), pairs.end(), less);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        volatile uint32_t sink = pairs[n / 2].first;
        (void) sink;
        return double(t1.tv_sec - t0.tv_sec) + 1e-9 * double(t1.tv_nsec - t0.tv_nsec);
    }

    // A fast stable CPU sort (LSD, 8-bit digits) used only to check multi-hundred-million-pair GPU
    // results in seconds; itself checked against std::stable_sort in tests/test_oracle.py.
    void glu_oracle_lsd_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, size_t num_steps)
    {
        uint32_t mask = steps_mask(num_steps);
        std::vector<uint32_t> k2(n), v2(n);
        uint32_t *sk = keys, *sv = vals, *dk = k2.data(), *dv = v2.data();
        for (int shift = 0; shift < 32; shift += 8)
        {
            uint32_t dmask = (mask >> shift) & 0xffu;
            if (dmask == 0)
                break;
            size_t hist[257] = {0};
            for (size_t i = 0; i < n; i++)
                hist[((sk[i] >> shift) & dmask) + 1]++;
            for (int d = 0; d < 256; d++)
                hist[d + 1] += hist[d];
            for (size_t i = 0; i < n; i++)
            {
                size_t p = hist[(sk[i] >> shift) & dmask]++;
                dk[p] = sk[i];
                dv[p] = sv[i];
            }
            std::swap(sk, dk);
            std::swap(sv, dv);
        }
        if (sk != keys)
        {
            std::memcpy(keys, sk, n * sizeof(uint32_t));
            std::memcpy(vals, sv, n * sizeof(uint32_t));
        }
    }

    // ---- (B) shader-faithful restatements ------------------------------------------------------------------------

    // data_type ids follow glu/data_types.hpp:8-22.  In place; result in element 0 (other elements are
    // clobbered with partials exactly as the reference does).
    int glu_oracle_reduce_glsl(void* data, size_t count, int data_type, int op)
    {
        if (!data || count == 0 || op < 0 || op > 3)
            return 1;
        switch (data_type)
        {
        case 0: reduce_tree<float>((float*) data, count, 1, op); break;
        case 1: reduce_tree<double>((double*) data, count, 1, op); break;
        case 2: reduce_tree<int32_t>((int32_t*) data, count, 1, op); break; // wraps like GLSL int
        case 3: reduce_tree<uint32_t>((uint32_t*) data, count, 1, op); break;
        case 4: reduce_tree<float>((float*) data, count, 2, op); break;
        case 5: reduce_tree<float>((float*) data, count, 4, op); break;
        case 6: reduce_tree<double>((double*) data, count, 2, op); break;
        case 7: reduce_tree<double>((double*) data, count, 4, op); break;
        case 8: reduce_tree<uint32_t>((uint32_t*) data, count, 2, op); break;
        case 9: reduce_tree<uint32_t>((uint32_t*) data, count, 4, op); break;
        case 10: reduce_tree<int32_t>((int32_t*) data, count, 2, op); break;
        case 11: reduce_tree<int32_t>((int32_t*) data, count, 4, op); break;
        default: return 1;
        }
        return 0;
    }

    // Returns non-zero where the reference would abort (glu/BlellochScan.hpp:132-135).
    int glu_oracle_blelloch_scan_glsl_u32(uint32_t* data, size_t count, size_t num_partitions)
    {
        if (!data || count == 0 || (count & (count - 1)) != 0 || num_partitions < 1)
            return 1;
        blelloch_scan<uint32_t>(data, count, num_partitions);
        return 0;
    }

    // In place on keys/vals (the result is copied back from the scratch when the reference would have
    // left it there); the return value is the buffer-pair index the reference would leave it in.
    int glu_oracle_radix_sort_glsl(uint32_t* keys, uint32_t* vals, size_t count, size_t num_steps)
    {
        std::vector<uint32_t> kb[2], vb[2];
        kb[0].assign(keys, keys + count);
        vb[0].assign(vals, vals + count);
        int where = radix_sort_glsl(kb, vb, count, num_steps);
        std::memcpy(keys, kb[where].data(), count * sizeof(uint32_t));
        std::memcpy(vals, vb[where].data(), count * sizeof(uint32_t));
        return where;
    }
}
