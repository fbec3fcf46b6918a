// glu_radix_sort_wide.cu — glu_radix_sort_wide(): 64-bit keys and payloads wider than 32 bits (SURVEY.md §8f row 3;
// the reference sorts uint32 keys with a mandatory uint32 value only, README.md:88-89, glu/RadixSort.hpp:273).
//
// Built ON the 32-bit onesweep sort instead of beside it: what is sorted is always a (32-bit key word, 32-bit payload)
// pair array — the shape the tuned kernels of glu_radix_sort.cu are written for.  A 64-bit key is two 32-bit "digits":
// a stable sort by the low word followed by a stable sort by the high word (LSD).
//   8-byte keys, no values   : the two words are each other's payload — sort (low, high), sort (high, low), zip.
//                              No random access at all: 25.0 Gkeys/s at 2^27 (B200).
//   8-byte keys, 4-byte value: the value rides through the same two sorts in two extra sorts on copies of the same
//                              sort keys (equal keys + stability = the same permutation): four coalesced sorts,
//                              13.0 Gpairs/s at 2^27 (two sorts + three random gathers measured 8.7).
//   any other combination    : the general path — sort (key word, element index), twice for 8-byte keys with a gather
//                              of the high words in between, and move the wide keys / values ONCE through the final
//                              permutation (gathers into scratch, copied back).  Payloads of 8 or 16 bytes ride along
//                              unchanged; the random gathers cost a 32-byte sector per element (8.3 Gpairs/s for
//                              8-byte keys + 8-byte values, 19.3 Gpairs/s for 4-byte keys + 16-byte values at 2^27).
// Stable (every sort is), ascending or descending (every sort complements its digits).  A native 8-pass onesweep over
// 12..16-byte pairs would move ~200 B per pair against ~300 B here; it is listed as future work in DESIGN.md §8.
#include "glu_common.cuh"

namespace glu_b200
{
    namespace
    {
        constexpr int k_wide_threads = 256;

        unsigned wide_grid(size_t n, int sms)
        {
            const size_t want = (n + k_wide_threads - 1) / k_wide_threads;
            const size_t cap = size_t(sms) * 16;
            return unsigned(want < 1 ? 1 : (want > cap ? cap : want));
        }

        // idx[i] = i; word[i] = low 32 bits of keys[i] (keys == nullptr: only the index array is written)
        __global__ void __launch_bounds__(k_wide_threads)
            wide_split_kernel(const uint64_t* __restrict__ keys, uint32_t* __restrict__ word, uint32_t* __restrict__ idx,
                              size_t n)
        {
            for (size_t i = size_t(blockIdx.x) * k_wide_threads + threadIdx.x; i < n; i += size_t(gridDim.x) * k_wide_threads)
            {
                idx[i] = uint32_t(i);
                if (keys)
                    word[i] = uint32_t(keys[i]);
            }
        }

        // low[i], high[i] = the two words of keys[i]
        __global__ void __launch_bounds__(k_wide_threads)
            wide_unzip_kernel(const uint64_t* __restrict__ keys, uint32_t* __restrict__ low, uint32_t* __restrict__ high,
                              size_t n)
        {
            for (size_t i = size_t(blockIdx.x) * k_wide_threads + threadIdx.x; i < n; i += size_t(gridDim.x) * k_wide_threads)
            {
                const uint64_t k = keys[i];
                low[i] = uint32_t(k);
                high[i] = uint32_t(k >> 32);
            }
        }

        // keys[i] = high[i] : low[i]
        __global__ void __launch_bounds__(k_wide_threads)
            wide_zip_kernel(const uint32_t* __restrict__ low, const uint32_t* __restrict__ high, uint64_t* __restrict__ keys,
                            size_t n)
        {
            for (size_t i = size_t(blockIdx.x) * k_wide_threads + threadIdx.x; i < n; i += size_t(gridDim.x) * k_wide_threads)
                keys[i] = (uint64_t(high[i]) << 32) | low[i];
        }

        // word[i] = high 32 bits of keys[idx[i]]
        __global__ void __launch_bounds__(k_wide_threads)
            wide_gather_high_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                    uint32_t* __restrict__ word, size_t n)
        {
            for (size_t i = size_t(blockIdx.x) * k_wide_threads + threadIdx.x; i < n; i += size_t(gridDim.x) * k_wide_threads)
                word[i] = uint32_t(keys[idx[i]] >> 32);
        }

        // out[i] = src[idx[i]] for 4-, 8- and 16-byte elements
        template<typename T>
        __global__ void __launch_bounds__(k_wide_threads)
            wide_gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ out, size_t n)
        {
            for (size_t i = size_t(blockIdx.x) * k_wide_threads + threadIdx.x; i < n; i += size_t(gridDim.x) * k_wide_threads)
                out[i] = src[idx[i]];
        }

        int launch_gather(const void* src, const uint32_t* idx, void* out, size_t elem_bytes, size_t n, int sms,
                          cudaStream_t s)
        {
            const unsigned grid = wide_grid(n, sms);
            switch (elem_bytes)
            {
            case 4:
                wide_gather_kernel<uint32_t><<<grid, k_wide_threads, 0, s>>>(static_cast<const uint32_t*>(src), idx,
                                                                            static_cast<uint32_t*>(out), n);
                break;
            case 8:
                wide_gather_kernel<uint64_t><<<grid, k_wide_threads, 0, s>>>(static_cast<const uint64_t*>(src), idx,
                                                                            static_cast<uint64_t*>(out), n);
                break;
            case 16:
                wide_gather_kernel<uint4><<<grid, k_wide_threads, 0, s>>>(static_cast<const uint4*>(src), idx,
                                                                         static_cast<uint4*>(out), n);
                break;
            default: return GLU_ERROR_INVALID_ARGUMENT;
            }
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }

        bool valid_widths(size_t key_bytes, size_t value_bytes)
        {
            return (key_bytes == 4 || key_bytes == 8) &&
                   (value_bytes == 0 || value_bytes == 4 || value_bytes == 8 || value_bytes == 16);
        }

        // the cases glu_radix_sort_u32_ex handles by itself
        bool is_narrow(size_t key_bytes, size_t value_bytes) { return key_bytes == 4 && value_bytes <= 4; }

        struct WideLayout
        {
            size_t off_idx, off_word, off_keys, off_vals, off_sort, sort_bytes, total;
        };

        WideLayout make_wide_layout(size_t count, size_t key_bytes, size_t value_bytes)
        {
            WideLayout l{};
            l.off_idx = 0;
            l.off_word = l.off_idx + align_up(count * sizeof(uint32_t), k_tmp_align);
            l.off_keys = l.off_word + (key_bytes == 8 ? align_up(count * sizeof(uint32_t), k_tmp_align) : 0);
            // (8-byte keys without values: [low words][high words][sort scratch], no permuted copy of the keys)
            l.off_vals = l.off_keys + (key_bytes == 8 && value_bytes != 0 ? align_up(count * key_bytes, k_tmp_align) : 0);
            l.off_sort = l.off_vals + align_up(count * value_bytes, k_tmp_align);
            l.sort_bytes = glu_radix_sort_u32_ex_tmp_bytes(count, 1);
            l.total = l.off_sort + l.sort_bytes;
            return l;
        }
    } // namespace
} // namespace glu_b200

using namespace glu_b200;

extern "C" size_t glu_radix_sort_wide_tmp_bytes(size_t count, size_t key_bytes, size_t value_bytes)
{
    if (!valid_widths(key_bytes, value_bytes))
        return 0;
    if (is_narrow(key_bytes, value_bytes))
        return glu_radix_sort_u32_ex_tmp_bytes(count, value_bytes != 0);
    if (count <= 1)
        return k_tmp_align;
    if (glu_radix_sort_u32_ex_tmp_bytes(count, 1) == 0) // count too large for the 32-bit sort underneath
        return 0;
    return make_wide_layout(count, key_bytes, value_bytes).total;
}

extern "C" int glu_radix_sort_wide(void* d_keys, size_t key_bytes, void* d_vals, size_t value_bytes, size_t count,
                                   int descending, void* d_tmp, size_t tmp_bytes, glu_stream_t stream)
{
    if (!d_keys || !valid_widths(key_bytes, value_bytes) || ((value_bytes != 0) != (d_vals != nullptr)))
        return GLU_ERROR_INVALID_ARGUMENT;
    if (is_narrow(key_bytes, value_bytes))
        return glu_radix_sort_u32_ex(static_cast<uint32_t*>(d_keys), static_cast<uint32_t*>(d_vals), count, 0, 32,
                                     descending, d_tmp, tmp_bytes, stream);
    if (count <= 1)
        return GLU_SUCCESS;
    if (glu_radix_sort_u32_ex_tmp_bytes(count, 1) == 0)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if (reinterpret_cast<uintptr_t>(d_keys) % key_bytes != 0 ||
        (d_vals && reinterpret_cast<uintptr_t>(d_vals) % value_bytes != 0))
        return GLU_ERROR_MISALIGNED;
    const WideLayout l = make_wide_layout(count, key_bytes, value_bytes);
    if (!d_tmp || tmp_bytes < l.total)
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    const int sms = current_sm_count();
    if (sms <= 0)
        return GLU_ERROR_CUDA;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char* tmp = static_cast<char*>(d_tmp);
    uint32_t* idx = reinterpret_cast<uint32_t*>(tmp + l.off_idx);
    void* sort_tmp = tmp + l.off_sort;
    const unsigned grid = wide_grid(count, sms);
    int rc = GLU_SUCCESS;

    if (key_bytes == 8 && value_bytes == 0)
    {
        uint32_t* low = idx; // the "index" array holds the low words here
        uint32_t* high = reinterpret_cast<uint32_t*>(tmp + l.off_word);
        wide_unzip_kernel<<<grid, k_wide_threads, 0, s>>>(static_cast<const uint64_t*>(d_keys), low, high, count);
        GLU_LAUNCH_CHECK();
        rc = glu_radix_sort_u32_ex(low, high, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        rc = glu_radix_sort_u32_ex(high, low, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        wide_zip_kernel<<<grid, k_wide_threads, 0, s>>>(low, high, static_cast<uint64_t*>(d_keys), count);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }
    if (key_bytes == 8 && value_bytes == 4)
    {
        // 8-byte keys with a 4-byte value: still no random access.  The value rides through the same two stable sorts
        // as the other key word, in two extra sorts on copies of the same sort keys (equal keys + stability = the same
        // permutation): four coalesced 68 B/pair sorts beat two sorts plus three random gathers (15.4 -> ~10 ms at 2^27).
        uint32_t* low = idx;
        uint32_t* high = reinterpret_cast<uint32_t*>(tmp + l.off_word);
        uint32_t* dup = reinterpret_cast<uint32_t*>(tmp + l.off_keys); // count * 8 bytes are reserved there
        uint32_t* vals = static_cast<uint32_t*>(d_vals);
        const size_t bytes = count * sizeof(uint32_t);
        wide_unzip_kernel<<<grid, k_wide_threads, 0, s>>>(static_cast<const uint64_t*>(d_keys), low, high, count);
        GLU_LAUNCH_CHECK();
        GLU_CUDA_TRY(cudaMemcpyAsync(dup, low, bytes, cudaMemcpyDeviceToDevice, s));
        rc = glu_radix_sort_u32_ex(low, high, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        rc = glu_radix_sort_u32_ex(dup, vals, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        GLU_CUDA_TRY(cudaMemcpyAsync(dup, high, bytes, cudaMemcpyDeviceToDevice, s));
        rc = glu_radix_sort_u32_ex(high, low, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        rc = glu_radix_sort_u32_ex(dup, vals, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        wide_zip_kernel<<<grid, k_wide_threads, 0, s>>>(low, high, static_cast<uint64_t*>(d_keys), count);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }
    if (key_bytes == 8)
    {
        const uint64_t* keys = static_cast<const uint64_t*>(d_keys);
        uint32_t* word = reinterpret_cast<uint32_t*>(tmp + l.off_word);
        void* keys_out = tmp + l.off_keys;
        wide_split_kernel<<<grid, k_wide_threads, 0, s>>>(keys, word, idx, count);
        GLU_LAUNCH_CHECK();
        rc = glu_radix_sort_u32_ex(word, idx, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        wide_gather_high_kernel<<<grid, k_wide_threads, 0, s>>>(keys, idx, word, count);
        GLU_LAUNCH_CHECK();
        rc = glu_radix_sort_u32_ex(word, idx, count, 0, 32, descending, sort_tmp, l.sort_bytes, stream);
        if (rc != GLU_SUCCESS)
            return rc;
        rc = launch_gather(d_keys, idx, keys_out, key_bytes, count, sms, s);
        if (rc != GLU_SUCCESS)
            return rc;
        GLU_CUDA_TRY(cudaMemcpyAsync(d_keys, keys_out, count * key_bytes, cudaMemcpyDeviceToDevice, s));
    }
    else
    {
        // 4-byte keys, 8- or 16-byte values: the keys themselves are sorted in place, carrying the index
        wide_split_kernel<<<grid, k_wide_threads, 0, s>>>(nullptr, nullptr, idx, count);
        GLU_LAUNCH_CHECK();
        rc = glu_radix_sort_u32_ex(static_cast<uint32_t*>(d_keys), idx, count, 0, 32, descending, sort_tmp, l.sort_bytes,
                                   stream);
        if (rc != GLU_SUCCESS)
            return rc;
    }
    if (value_bytes != 0)
    {
        void* vals_out = tmp + l.off_vals;
        rc = launch_gather(d_vals, idx, vals_out, value_bytes, count, sms, s);
        if (rc != GLU_SUCCESS)
            return rc;
        GLU_CUDA_TRY(cudaMemcpyAsync(d_vals, vals_out, count * value_bytes, cudaMemcpyDeviceToDevice, s));
    }
    return GLU_SUCCESS;
}
