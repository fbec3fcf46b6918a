// glu_common.cuh — shared host/device helpers for the sm_100a kernels behind include/glu_b200.h.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

#include "glu_b200.h"

namespace glu_b200
{
    // ---------------------------------------------------------------------------------------------- host side

    extern std::atomic<uint64_t> g_kernel_launches;
    extern thread_local cudaError_t t_last_cuda_error;

    inline int cuda_fail(cudaError_t e)
    {
        t_last_cuda_error = e;
        return GLU_ERROR_CUDA;
    }

#define GLU_CUDA_TRY(expr)                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return ::glu_b200::cuda_fail(e__);                                                                         \
    } while (0)

    // Call right after a <<<>>> launch.
#define GLU_LAUNCH_CHECK()                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        ::glu_b200::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                                         \
        GLU_CUDA_TRY(cudaGetLastError());                                                                              \
    } while (0)

    // Optional event bracketing of a launch (glu_profile_enable): construct before <<<>>>, destroyed after.
    extern std::atomic<int> g_profile_on;
    cudaEvent_t profile_begin(int kernel_id, cudaStream_t s); // returns the span's stop event
    void profile_end(cudaEvent_t stop, cudaStream_t s);
    struct ScopedKernelProfile
    {
        cudaEvent_t stop = nullptr;
        cudaStream_t s;
        ScopedKernelProfile(int kernel_id, cudaStream_t stream) : s(stream)
        {
            if (g_profile_on.load(std::memory_order_relaxed) != 0)
                stop = profile_begin(kernel_id, s);
        }
        ~ScopedKernelProfile()
        {
            if (stop)
                profile_end(stop, s);
        }
    };

    constexpr size_t k_tmp_align = 256;
    inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

    // SM count of the current device, cached per device (grid sizing in multiples of the SM count).
    int current_sm_count();

    struct DataTypeInfo
    {
        int scalar; // 0 = f32, 1 = f64, 2 = i32, 3 = u32
        int ncomp;  // 1, 2, 4
        size_t scalar_size;
    };
    // glu/data_types.hpp:8-22; returns false for an invalid id.
    bool data_type_info(int data_type, DataTypeInfo* out);

    // ---------------------------------------------------------------------------------------------- device side

    constexpr unsigned k_full_mask = 0xffffffffu;

#ifdef __CUDACC__
    __device__ __forceinline__ unsigned lane_id()
    {
        unsigned r;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
        return r;
    }
    __device__ __forceinline__ unsigned lanemask_lt()
    {
        unsigned r;
        asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
        return r;
    }
    __device__ __forceinline__ unsigned lanemask_le()
    {
        unsigned r;
        asm volatile("mov.u32 %0, %%lanemask_le;" : "=r"(r));
        return r;
    }

    // Relaxed / acquire / release accesses at gpu scope for the decoupled look-back protocols.
    __device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
    {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    __device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
    {
        uint32_t v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
    {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    __device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
    {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    __device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v)
    {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p)
    {
        uint64_t v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
    }

    // Streaming (evict-first) 128-bit global load / store: data that is touched exactly once.  The loads are
    // deliberately not `volatile`: every user consumes the value before it stores to the same location, so
    // ordering follows from data dependence and ptxas is free to keep many loads in flight.
    __device__ __forceinline__ uint4 ld_stream_v4(const void* p)
    {
        uint4 r;
        asm("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                     : "l"(p));
        return r;
    }
    __device__ __forceinline__ void st_stream_v4(void* p, uint4 v)
    {
        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
    }
    __device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
    {
        uint32_t r;
        asm("ld.global.cs.u32 %0, [%1];" : "=r"(r) : "l"(p));
        return r;
    }

    // ---- TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier completion ----------------------------
    __device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

    __device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t arrivals)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    }
    // makes the initialised barrier visible to the async proxy (follow with a CTA barrier)
    __device__ __forceinline__ void mbarrier_init_fence()
    {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void mbarrier_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                     : "memory");
    }
    __device__ __forceinline__ void mbarrier_arrive(uint64_t* bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    __device__ __forceinline__ void named_barrier_sync(int id, int threads)
    {
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
    }
    __device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity)
    {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "WAIT_%=:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                     "@p bra DONE_%=;\n"
                     "bra WAIT_%=;\n"
                     "DONE_%=:\n"
                     "}" ::"r"(smem_u32(bar)),
                     "r"(parity)
                     : "memory");
    }
    // orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
    __device__ __forceinline__ void fence_proxy_async_smem()
    {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __device__ __forceinline__ uint64_t l2_policy_evict_first()
    {
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        return policy;
    }
    // global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled
    // on `bar` as a transaction-byte count.
    __device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                uint64_t policy)
    {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                     "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                     "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                     : "memory");
    }
    // Asks L2 to fetch `bytes` (a multiple of 16, 16-byte aligned address) ahead of a later bulk copy; no completion
    // is signalled and nothing lands in shared memory (SASS: UBLKPF).
    __device__ __forceinline__ void tma_prefetch_l2_1d(const void* gmem_src, uint32_t bytes)
    {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
    }
#endif // __CUDACC__
} // namespace glu_b200
