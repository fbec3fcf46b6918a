// glu_reduce.cu — glu_reduce(): the B200 replacement for glu::Reduce::operator() (glu/Reduce.hpp:111-135).
//
// The reference walks a radix-32 tree with one dispatch per level (one element per thread, one
// subgroup op per level, ceil(log32 N) dispatches, strided re-reads of the partials).  Here the whole
// reduction is ONE kernel launch:
//   * a grid of (SM count x k_blocks_per_sm) CTAs strides over the buffer with 128-bit streaming
//     loads, k_unroll loads in flight per thread (HBM-bound: 4 B of traffic per 32-bit element);
//   * per-thread accumulators -> __reduce_*_sync (32-bit integer sum/min/max) or a shuffle tree
//     (float, double, product) -> shared memory -> one partial per CTA in temporary storage;
//   * the last CTA to finish (atomic ticket) folds the partials in a fixed order and stores the
//     result to element 0, so the result is deterministic run to run for floating types too.
// Vector types (vec2/vec4/...) are reduced component-wise: a 16-byte load holds 16/sizeof(scalar)
// scalars whose component index depends only on the thread, never on the loop iteration.
#include "glu_common.cuh"

namespace glu_b200
{
    namespace
    {
        constexpr int k_threads = 512;
        constexpr int k_blocks_per_sm = 2;
        constexpr int k_unroll = 8;

        template<typename S, int OP> struct Operator;
        template<typename S> struct Operator<S, GLU_REDUCE_OPERATOR_SUM>
        {
            static __device__ __forceinline__ S identity() { return S(0); }
            static __device__ __forceinline__ S apply(S a, S b) { return a + b; }
        };
        template<typename S> struct Operator<S, GLU_REDUCE_OPERATOR_MUL>
        {
            static __device__ __forceinline__ S identity() { return S(1); }
            static __device__ __forceinline__ S apply(S a, S b) { return a * b; }
        };
        template<> struct Operator<uint32_t, GLU_REDUCE_OPERATOR_MIN>
        {
            static __device__ __forceinline__ uint32_t identity() { return 0xffffffffu; }
            static __device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b) { return min(a, b); }
        };
        template<> struct Operator<uint32_t, GLU_REDUCE_OPERATOR_MAX>
        {
            static __device__ __forceinline__ uint32_t identity() { return 0u; }
            static __device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b) { return max(a, b); }
        };
        template<> struct Operator<int32_t, GLU_REDUCE_OPERATOR_MIN>
        {
            static __device__ __forceinline__ int32_t identity() { return 0x7fffffff; }
            static __device__ __forceinline__ int32_t apply(int32_t a, int32_t b) { return min(a, b); }
        };
        template<> struct Operator<int32_t, GLU_REDUCE_OPERATOR_MAX>
        {
            static __device__ __forceinline__ int32_t identity() { return int32_t(0x80000000u); }
            static __device__ __forceinline__ int32_t apply(int32_t a, int32_t b) { return max(a, b); }
        };
        template<> struct Operator<float, GLU_REDUCE_OPERATOR_MIN>
        {
            static __device__ __forceinline__ float identity() { return __int_as_float(0x7f800000); }
            static __device__ __forceinline__ float apply(float a, float b) { return fminf(a, b); }
        };
        template<> struct Operator<float, GLU_REDUCE_OPERATOR_MAX>
        {
            static __device__ __forceinline__ float identity() { return __int_as_float(0xff800000); }
            static __device__ __forceinline__ float apply(float a, float b) { return fmaxf(a, b); }
        };
        template<> struct Operator<double, GLU_REDUCE_OPERATOR_MIN>
        {
            static __device__ __forceinline__ double identity() { return __longlong_as_double(0x7ff0000000000000ll); }
            static __device__ __forceinline__ double apply(double a, double b) { return fmin(a, b); }
        };
        template<> struct Operator<double, GLU_REDUCE_OPERATOR_MAX>
        {
            static __device__ __forceinline__ double identity() { return __longlong_as_double(0xfff0000000000000ll); }
            static __device__ __forceinline__ double apply(double a, double b) { return fmax(a, b); }
        };

        // int32 sum/product wrap like GLSL ints: do them on the unsigned representation.
        template<> struct Operator<int32_t, GLU_REDUCE_OPERATOR_SUM>
        {
            static __device__ __forceinline__ int32_t identity() { return 0; }
            static __device__ __forceinline__ int32_t apply(int32_t a, int32_t b)
            {
                return int32_t(uint32_t(a) + uint32_t(b));
            }
        };
        template<> struct Operator<int32_t, GLU_REDUCE_OPERATOR_MUL>
        {
            static __device__ __forceinline__ int32_t identity() { return 1; }
            static __device__ __forceinline__ int32_t apply(int32_t a, int32_t b)
            {
                return int32_t(uint32_t(a) * uint32_t(b));
            }
        };

        // ---- warp reduction: redux.sync for 32-bit integer sum/min/max, shuffle tree otherwise ----
        template<typename S, int OP> __device__ __forceinline__ S warp_reduce(S v)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v = Operator<S, OP>::apply(v, __shfl_xor_sync(k_full_mask, v, o));
            return v;
        }
        template<> __device__ __forceinline__ uint32_t warp_reduce<uint32_t, GLU_REDUCE_OPERATOR_SUM>(uint32_t v)
        {
            return __reduce_add_sync(k_full_mask, v);
        }
        template<> __device__ __forceinline__ uint32_t warp_reduce<uint32_t, GLU_REDUCE_OPERATOR_MIN>(uint32_t v)
        {
            return __reduce_min_sync(k_full_mask, v);
        }
        template<> __device__ __forceinline__ uint32_t warp_reduce<uint32_t, GLU_REDUCE_OPERATOR_MAX>(uint32_t v)
        {
            return __reduce_max_sync(k_full_mask, v);
        }
        template<> __device__ __forceinline__ int32_t warp_reduce<int32_t, GLU_REDUCE_OPERATOR_SUM>(int32_t v)
        {
            return int32_t(__reduce_add_sync(k_full_mask, uint32_t(v)));
        }
        template<> __device__ __forceinline__ int32_t warp_reduce<int32_t, GLU_REDUCE_OPERATOR_MIN>(int32_t v)
        {
            return __reduce_min_sync(k_full_mask, v);
        }
        template<> __device__ __forceinline__ int32_t warp_reduce<int32_t, GLU_REDUCE_OPERATOR_MAX>(int32_t v)
        {
            return __reduce_max_sync(k_full_mask, v);
        }

        template<typename S> struct Unit // one 16-byte load seen as scalars
        {
            static constexpr int L = 16 / int(sizeof(S));
            union
            {
                uint4 raw;
                S s[L];
            };
        };

        // Block-wide fold of NCOMP per-thread values; the result is valid in thread 0.
        template<typename S, int NCOMP, int OP, int THREADS>
        __device__ __forceinline__ void block_reduce(S (&r)[NCOMP], S (*s_warp)[NCOMP])
        {
            constexpr int WARPS = THREADS / 32;
            const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
            for (int c = 0; c < NCOMP; c++)
                r[c] = warp_reduce<S, OP>(r[c]);
            if (lane == 0)
            {
#pragma unroll
                for (int c = 0; c < NCOMP; c++)
                    s_warp[warp][c] = r[c];
            }
            __syncthreads();
            if (warp == 0)
            {
#pragma unroll
                for (int c = 0; c < NCOMP; c++)
                {
                    S v = lane < WARPS ? s_warp[lane][c] : Operator<S, OP>::identity();
                    r[c] = warp_reduce<S, OP>(v);
                }
            }
            __syncthreads(); // s_warp may be reused by the caller
        }

        // data      : base pointer of the buffer (scalar view)
        // out       : where the NCOMP result scalars go (data itself for the in-place glu_reduce)
        // head      : scalars before the first 16-byte aligned address (handled by CTA 0)
        // n_units   : number of whole 16-byte units after the head
        // n_scalars : count * NCOMP
        template<typename S, int NCOMP, int OP, int THREADS, int UNROLL>
        __global__ void __launch_bounds__(THREADS)
            reduce_kernel(const S* data, S* out, // no __restrict__: out == data for the in-place glu_reduce
                          size_t n_scalars, unsigned head, size_t n_units, S* partials, unsigned* ticket)
        {
            using Opr = Operator<S, OP>;
            constexpr int L = Unit<S>::L;
            constexpr int WARPS = THREADS / 32;
            __shared__ S s_warp[WARPS][NCOMP];
            __shared__ bool s_is_last;

            const uint4* body = reinterpret_cast<const uint4*>(data + head);
            const size_t stride = size_t(gridDim.x) * THREADS; // even => a thread's component phase is fixed
            size_t v = size_t(blockIdx.x) * THREADS + threadIdx.x;
            const unsigned phase = unsigned((head + v * L) & (NCOMP - 1)); // component of acc[0]

            S acc[L];
#pragma unroll
            for (int k = 0; k < L; k++)
                acc[k] = Opr::identity();

            for (; v + size_t(UNROLL - 1) * stride < n_units; v += size_t(UNROLL) * stride)
            {
                Unit<S> u[UNROLL];
#pragma unroll
                for (int j = 0; j < UNROLL; j++)
                    u[j].raw = ld_stream_v4(body + v + size_t(j) * stride);
#pragma unroll
                for (int j = 0; j < UNROLL; j++)
#pragma unroll
                    for (int k = 0; k < L; k++)
                        acc[k] = Opr::apply(acc[k], u[j].s[k]);
            }
            for (; v < n_units; v += stride)
            {
                Unit<S> u;
                u.raw = ld_stream_v4(body + v);
#pragma unroll
                for (int k = 0; k < L; k++)
                    acc[k] = Opr::apply(acc[k], u.s[k]);
            }

            // fold the L lanes of the 16-byte unit onto the NCOMP components
            S r[NCOMP];
#pragma unroll
            for (int c = 0; c < NCOMP; c++)
            {
                r[c] = Opr::identity();
#pragma unroll
                for (int k = 0; k < L; k++)
                    if (((phase + k) & (NCOMP - 1)) == unsigned(c))
                        r[c] = Opr::apply(r[c], acc[k]);
            }

            // unaligned head / sub-unit tail scalars (at most L-1 each): CTA 0
            if (blockIdx.x == 0)
            {
                const size_t tail_begin = size_t(head) + n_units * L;
                size_t s = threadIdx.x < head ? size_t(threadIdx.x) : tail_begin + (threadIdx.x - head);
                if (s < n_scalars && (threadIdx.x < head || s >= tail_begin))
                {
                    S x = data[s];
#pragma unroll
                    for (int c = 0; c < NCOMP; c++)
                        if ((s & (NCOMP - 1)) == size_t(c))
                            r[c] = Opr::apply(r[c], x);
                }
            }

            block_reduce<S, NCOMP, OP, THREADS>(r, s_warp);

            if (gridDim.x == 1)
            {
                if (threadIdx.x == 0)
                {
#pragma unroll
                    for (int c = 0; c < NCOMP; c++)
                        out[c] = r[c];
                }
                return;
            }

            if (threadIdx.x == 0)
            {
#pragma unroll
                for (int c = 0; c < NCOMP; c++)
                    partials[size_t(blockIdx.x) * NCOMP + c] = r[c];
                __threadfence();
                unsigned t = atomicAdd(ticket, 1u);
                s_is_last = (t == gridDim.x - 1);
            }
            __syncthreads();
            if (!s_is_last)
                return;
            __threadfence();

            // last CTA: fixed-order fold of the per-CTA partials (deterministic)
#pragma unroll
            for (int c = 0; c < NCOMP; c++)
                r[c] = Opr::identity();
            for (unsigned b = threadIdx.x; b < gridDim.x; b += THREADS)
            {
#pragma unroll
                for (int c = 0; c < NCOMP; c++)
                    r[c] = Opr::apply(r[c], __ldcg(&partials[size_t(b) * NCOMP + c]));
            }
            block_reduce<S, NCOMP, OP, THREADS>(r, s_warp);
            if (threadIdx.x == 0)
            {
#pragma unroll
                for (int c = 0; c < NCOMP; c++)
                    out[c] = r[c];
            }
        }

        inline int reduce_grid(size_t n_units)
        {
            size_t per_block = size_t(k_threads) * k_unroll;
            size_t want = (n_units + per_block - 1) / per_block;
            size_t cap = size_t(current_sm_count()) * k_blocks_per_sm;
            if (want < 1)
                want = 1;
            return int(want < cap ? want : cap);
        }

        template<typename S, int NCOMP, int OP>
        int launch_reduce(const void* d_data, void* d_out, size_t count, void* d_tmp, cudaStream_t stream)
        {
            constexpr int L = Unit<S>::L;
            const size_t n_scalars = count * NCOMP;
            const uintptr_t addr = reinterpret_cast<uintptr_t>(d_data);
            unsigned head = unsigned(((16 - (addr & 15)) & 15) / sizeof(S));
            if (head > n_scalars)
                head = unsigned(n_scalars);
            const size_t n_units = (n_scalars - head) / L;
            const int grid = reduce_grid(n_units);
            unsigned* ticket = reinterpret_cast<unsigned*>(d_tmp);
            S* partials = reinterpret_cast<S*>(reinterpret_cast<char*>(d_tmp) + k_tmp_align);
            if (grid > 1)
                GLU_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), stream));
            ScopedKernelProfile prof(GLU_KERNEL_REDUCE, stream);
            reduce_kernel<S, NCOMP, OP, k_threads, k_unroll>
                <<<grid, k_threads, 0, stream>>>(reinterpret_cast<const S*>(d_data), reinterpret_cast<S*>(d_out), n_scalars, head,
                                                 n_units, partials, ticket);
            GLU_LAUNCH_CHECK();
            return GLU_SUCCESS;
        }

        template<typename S, int NCOMP> int dispatch_op(const void* d, void* o, size_t n, int op, void* tmp, cudaStream_t s)
        {
            switch (op)
            {
            case GLU_REDUCE_OPERATOR_SUM: return launch_reduce<S, NCOMP, GLU_REDUCE_OPERATOR_SUM>(d, o, n, tmp, s);
            case GLU_REDUCE_OPERATOR_MUL: return launch_reduce<S, NCOMP, GLU_REDUCE_OPERATOR_MUL>(d, o, n, tmp, s);
            case GLU_REDUCE_OPERATOR_MIN: return launch_reduce<S, NCOMP, GLU_REDUCE_OPERATOR_MIN>(d, o, n, tmp, s);
            case GLU_REDUCE_OPERATOR_MAX: return launch_reduce<S, NCOMP, GLU_REDUCE_OPERATOR_MAX>(d, o, n, tmp, s);
            default: return GLU_ERROR_INVALID_OPERATOR;
            }
        }

        template<typename S> int dispatch_ncomp(const void* d, void* o, size_t n, int ncomp, int op, void* tmp, cudaStream_t s)
        {
            switch (ncomp)
            {
            case 1: return dispatch_op<S, 1>(d, o, n, op, tmp, s);
            case 2: return dispatch_op<S, 2>(d, o, n, op, tmp, s);
            default: return dispatch_op<S, 4>(d, o, n, op, tmp, s);
            }
        }
    } // namespace
} // namespace glu_b200

using namespace glu_b200;

extern "C" size_t glu_reduce_tmp_bytes(size_t count, int data_type)
{
    (void) count;
    DataTypeInfo info;
    if (!data_type_info(data_type, &info))
        return 0;
    // ticket (one aligned slot) + one partial per CTA of the largest grid ever launched
    size_t max_grid = 1024 * size_t(k_blocks_per_sm); // >= SM count x k_blocks_per_sm on any device
    return k_tmp_align + align_up(max_grid * info.ncomp * info.scalar_size, k_tmp_align);
}

extern "C" int glu_reduce(void* d_data, size_t count, int data_type, int op, void* d_tmp, size_t tmp_bytes,
                          glu_stream_t stream)
{
    return glu_reduce_into(d_data, count, data_type, op, d_data, d_tmp, tmp_bytes, stream);
}

extern "C" int glu_reduce_into(const void* d_data, size_t count, int data_type, int op, void* d_result, void* d_tmp,
                               size_t tmp_bytes, glu_stream_t stream)
{
    DataTypeInfo info;
    if (!data_type_info(data_type, &info))
        return GLU_ERROR_INVALID_DATA_TYPE;
    if (op < GLU_REDUCE_OPERATOR_SUM || op > GLU_REDUCE_OPERATOR_MAX)
        return GLU_ERROR_INVALID_OPERATOR;
    if (!d_data || !d_result || count == 0) // glu/Reduce.hpp:113-114
        return GLU_ERROR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_data) % info.scalar_size != 0 || reinterpret_cast<uintptr_t>(d_result) % info.scalar_size != 0)
        return GLU_ERROR_MISALIGNED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (count == 1)
    {
        if (d_result != d_data)
            GLU_CUDA_TRY(cudaMemcpyAsync(d_result, d_data, info.scalar_size * info.ncomp, cudaMemcpyDeviceToDevice, s));
        return GLU_SUCCESS;
    }
    if (!d_tmp || tmp_bytes < glu_reduce_tmp_bytes(count, data_type))
        return GLU_ERROR_TMP_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(d_tmp) % k_tmp_align != 0)
        return GLU_ERROR_MISALIGNED;
    if (current_sm_count() <= 0)
        return GLU_ERROR_CUDA;
    switch (info.scalar)
    {
    case 0: return dispatch_ncomp<float>(d_data, d_result, count, info.ncomp, op, d_tmp, s);
    case 1: return dispatch_ncomp<double>(d_data, d_result, count, info.ncomp, op, d_tmp, s);
    case 2: return dispatch_ncomp<int32_t>(d_data, d_result, count, info.ncomp, op, d_tmp, s);
    default: return dispatch_ncomp<uint32_t>(d_data, d_result, count, info.ncomp, op, d_tmp, s);
    }
}
