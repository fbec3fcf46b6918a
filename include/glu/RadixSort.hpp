// glu/RadixSort.hpp — glu::RadixSort for CUDA device buffers (reference: glu/RadixSort.hpp:186-354).
//
//     glu::RadixSort radix_sort;
//     radix_sort.prepare_internal_buffers(count);          // optional pre-sizing, as the reference's benchmark does
//     radix_sort(d_keys, d_vals, count);                   // stable, ascending, in place
//
// Keys and values are two separate dense uint32 arrays, as in the reference.  num_steps keeps the
// reference's meaning (number of 4-bit steps, i.e. only the low 4*num_steps key bits take part; 0 = all).
// Deviation: the result always ends up in the caller's buffers (the reference leaves an odd-num_steps
// result in its internal scratch).
#ifndef GLU_B200_RADIXSORT_HPP
#define GLU_B200_RADIXSORT_HPP

#include "BlellochScan.hpp"
#include "device_utils.hpp"

namespace glu
{
    class RadixSort
    {
    private:
        /// Ping-pong key/value arrays, digit histograms, per-tile look-back words (glu_radix_sort_u32kv_tmp_bytes).
        DeviceBuffer m_tmp;
        glu_stream_t m_stream = nullptr;

    public:
        explicit RadixSort() = default;
        ~RadixSort() = default;

        void set_stream(glu_stream_t stream) { m_stream = stream; }

        void prepare_internal_buffers(size_t count)
        {
            const size_t need = glu_radix_sort_u32kv_tmp_bytes(count);
            GLU_CHECK_ARGUMENT(need != 0, "RadixSort: count %zu is too large", count);
            if (m_tmp.size() < need)
            {
                m_tmp.resize(need, false);
#ifdef GLU_VERBOSE
                std::printf("[RadixSort] Scratch buffer reallocated to: %zu\n", need);
#endif
            }
        }

        void operator()(DevicePtr key_buffer, DevicePtr val_buffer, size_t count, size_t num_steps = 0)
        {
            GLU_CHECK_ARGUMENT(key_buffer, "Invalid key buffer");
            GLU_CHECK_ARGUMENT(val_buffer, "Invalid value buffer");
            if (count <= 1)
                return;
            prepare_internal_buffers(count);
            GLU_CHECK_STATUS(glu_radix_sort_u32kv(static_cast<uint32_t*>(key_buffer), static_cast<uint32_t*>(val_buffer),
                                                  count, num_steps, m_tmp.handle(), m_tmp.size(), m_stream));
        }

        /// Beyond the reference (README.md:88-89 lists the mandatory value buffer as a limitation):
        /// val_buffer may be null (key-only sort), only key bits [begin_bit, end_bit) take part, descending puts the
        /// largest key first.  Always stable.  glu_radix_sort_u32_ex.
        void sort_ex(DevicePtr key_buffer, DevicePtr val_buffer, size_t count, unsigned begin_bit = 0, unsigned end_bit = 32,
                     bool descending = false)
        {
            GLU_CHECK_ARGUMENT(key_buffer, "Invalid key buffer");
            GLU_CHECK_ARGUMENT(begin_bit <= end_bit && end_bit <= 32, "RadixSort: need begin_bit <= end_bit <= 32");
            if (count <= 1 || begin_bit == end_bit)
                return;
            const size_t need = glu_radix_sort_u32_ex_tmp_bytes(count, val_buffer ? 1 : 0);
            GLU_CHECK_ARGUMENT(need != 0, "RadixSort: count %zu is too large", count);
            if (m_tmp.size() < need)
                m_tmp.resize(need, false);
            GLU_CHECK_STATUS(glu_radix_sort_u32_ex(static_cast<uint32_t*>(key_buffer), static_cast<uint32_t*>(val_buffer),
                                                   count, begin_bit, end_bit, descending ? 1 : 0, m_tmp.handle(),
                                                   m_tmp.size(), m_stream));
        }

        /// 64-bit keys and wide payloads (glu_radix_sort_wide): key_bytes 4 or 8, value_bytes 0 (val_buffer null), 4, 8
        /// or 16.  Stable, in place.
        void sort_wide(DevicePtr key_buffer, size_t key_bytes, DevicePtr val_buffer, size_t value_bytes, size_t count,
                       bool descending = false)
        {
            GLU_CHECK_ARGUMENT(key_buffer, "Invalid key buffer");
            GLU_CHECK_ARGUMENT((val_buffer != nullptr) == (value_bytes != 0), "Invalid value buffer / value_bytes");
            if (count <= 1)
                return;
            const size_t need = glu_radix_sort_wide_tmp_bytes(count, key_bytes, value_bytes);
            GLU_CHECK_ARGUMENT(need != 0, "RadixSort: unsupported element widths (%zu, %zu) or count %zu too large",
                               key_bytes, value_bytes, count);
            if (m_tmp.size() < need)
                m_tmp.resize(need, false);
            GLU_CHECK_STATUS(glu_radix_sort_wide(key_buffer, key_bytes, val_buffer, value_bytes, count, descending ? 1 : 0,
                                                 m_tmp.handle(), m_tmp.size(), m_stream));
        }

        /// Many independent stable sorts by key bits [begin_bit, end_bit) in one set of launches
        /// (glu_radix_sort_u32kv_segmented; the local step of the multi-GPU sort).  Input in the A arrays, segment s —
        /// seg_count_buffer[s] pairs, a device array — starting at element first_tile[s] * segment_tile() with
        /// first_tile = exclusive scan of ceil(count / tile); output compact.  Returns true when the result is in the
        /// B arrays (odd number of 8-bit passes), false when it is in the A arrays.
        static size_t segment_tile() { return glu_radix_sort_segment_tile(); }
        bool sort_segmented(DevicePtr keys_a, DevicePtr vals_a, DevicePtr keys_b, DevicePtr vals_b,
                            DevicePtr seg_count_buffer, size_t num_segments, size_t max_tiles, unsigned begin_bit = 0,
                            unsigned end_bit = 32)
        {
            GLU_CHECK_ARGUMENT(keys_a && vals_a && keys_b && vals_b && seg_count_buffer, "Invalid buffer");
            const size_t need = glu_radix_sort_u32kv_segmented_tmp_bytes(max_tiles);
            GLU_CHECK_ARGUMENT(need != 0, "RadixSort: %zu tiles are too many", max_tiles);
            if (m_tmp.size() < need)
                m_tmp.resize(need, false);
            int in_b = 0;
            GLU_CHECK_STATUS(glu_radix_sort_u32kv_segmented(
                static_cast<uint32_t*>(keys_a), static_cast<uint32_t*>(vals_a), static_cast<uint32_t*>(keys_b),
                static_cast<uint32_t*>(vals_b), static_cast<const uint32_t*>(seg_count_buffer), num_segments, max_tiles,
                begin_bit, end_bit, m_tmp.handle(), m_tmp.size(), m_stream, &in_b));
            return in_b != 0;
        }

        /// The same with the input given as RUNS (glu_radix_sort_u32kv_segmented_runs): `runs_buffer` holds 5 rows of
        /// (num_runs + 1) uint32 — first tile of every run in the first pass's own numbering (+ the total), the tile of
        /// the A arrays it starts at, its count, its segment, the first tile of its segment.  What the multi-GPU sort
        /// calls after its copy-engine all-to-all (a bucket is spread over one chunk per source rank).
        bool sort_segmented(DevicePtr keys_a, DevicePtr vals_a, DevicePtr keys_b, DevicePtr vals_b,
                            DevicePtr seg_count_buffer, size_t num_segments, size_t max_tiles, unsigned begin_bit,
                            unsigned end_bit, DevicePtr runs_buffer, size_t num_runs)
        {
            GLU_CHECK_ARGUMENT(keys_a && vals_a && keys_b && vals_b && seg_count_buffer && runs_buffer, "Invalid buffer");
            const size_t need = glu_radix_sort_u32kv_segmented_tmp_bytes(max_tiles);
            GLU_CHECK_ARGUMENT(need != 0, "RadixSort: %zu tiles are too many", max_tiles);
            if (m_tmp.size() < need)
                m_tmp.resize(need, false);
            int in_b = 0;
            GLU_CHECK_STATUS(glu_radix_sort_u32kv_segmented_runs(
                static_cast<uint32_t*>(keys_a), static_cast<uint32_t*>(vals_a), static_cast<uint32_t*>(keys_b),
                static_cast<uint32_t*>(vals_b), static_cast<const uint32_t*>(seg_count_buffer), num_segments, max_tiles,
                begin_bit, end_bit, static_cast<const uint32_t*>(runs_buffer), num_runs, m_tmp.handle(), m_tmp.size(),
                m_stream, &in_b));
            return in_b != 0;
        }
    };
} // namespace glu

#endif // GLU_B200_RADIXSORT_HPP
