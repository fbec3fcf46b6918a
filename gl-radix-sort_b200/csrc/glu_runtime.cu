// glu_runtime.cu — the non-kernel part of the C ABI: status strings, data-type table, device memory /
// stream / event plumbing (the ShaderStorageBuffer + measure_gl_elapsed_time roles of
// glu/gl_utils.hpp:146-265) and the host-buffer entry points used for end-to-end measurements.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "glu_common.cuh"

namespace glu_b200
{
    std::atomic<uint64_t> g_kernel_launches{0};
    thread_local cudaError_t t_last_cuda_error = cudaSuccess;

    std::atomic<int> g_profile_on{0};
    namespace
    {
        struct ProfileSpan
        {
            int id;
            cudaEvent_t start, stop;
        };
        std::mutex g_profile_mutex;
        std::vector<ProfileSpan> g_profile_spans;
        std::vector<cudaEvent_t> g_profile_free_events;

        cudaEvent_t profile_event()
        {
            if (!g_profile_free_events.empty())
            {
                cudaEvent_t e = g_profile_free_events.back();
                g_profile_free_events.pop_back();
                return e;
            }
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            return e;
        }
    } // namespace

    // returns the span's stop event: the caller records it after its launch (ScopedKernelProfile), so spans opened
    // concurrently on two streams — or by two threads — can never be paired with each other's launches
    cudaEvent_t profile_begin(int kernel_id, cudaStream_t s)
    {
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        ProfileSpan span{kernel_id, profile_event(), profile_event()};
        cudaEventRecord(span.start, s);
        g_profile_spans.push_back(span);
        return span.stop;
    }

    void profile_end(cudaEvent_t stop, cudaStream_t s)
    {
        if (stop)
            cudaEventRecord(stop, s);
    }

    int current_sm_count()
    {
        static std::atomic<int> cache[64];
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        {
            t_last_cuda_error = cudaGetLastError();
            return 0;
        }
        int v = cache[dev].load(std::memory_order_relaxed);
        if (v == 0)
        {
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            {
                t_last_cuda_error = cudaGetLastError();
                return 0;
            }
            cache[dev].store(v, std::memory_order_relaxed);
        }
        return v;
    }

    bool data_type_info(int data_type, DataTypeInfo* out)
    {
        // {scalar kind, components}; scalar kind: 0 f32, 1 f64, 2 i32, 3 u32  (glu/data_types.hpp:8-22)
        static const int table[12][2] = {{0, 1}, {1, 1}, {2, 1}, {3, 1}, {0, 2}, {0, 4},
                                         {1, 2}, {1, 4}, {3, 2}, {3, 4}, {2, 2}, {2, 4}};
        if (data_type < 0 || data_type > 11)
            return false;
        out->scalar = table[data_type][0];
        out->ncomp = table[data_type][1];
        out->scalar_size = out->scalar == 1 ? 8 : 4;
        return true;
    }
} // namespace glu_b200

using namespace glu_b200;

namespace glu_b200
{
    __global__ void fill_u32_kernel(uint32_t* dst, uint32_t value, size_t count)
    {
        for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
            dst[i] = value;
    }

    // Stream-ordered signalling between GPUs (the multi-GPU sort's device-side barrier without a collective): a rank
    // stores an epoch number into one word of every peer's flag array AFTER its copies into that peer (same stream), a
    // peer's sorting stream waits until all its words have reached the epoch.
    struct PeerFlags
    {
        uint32_t* ptr[16];
    };

    __global__ void signal_peers_kernel(PeerFlags flags, int count, uint32_t value)
    {
        if (int(threadIdx.x) < count && flags.ptr[threadIdx.x])
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.ptr[threadIdx.x]), "r"(value) : "memory");
    }

    // thread i waits for flags[i] >= value (wrap-around safe), i != skip.  Bounded: a peer that never signals (it
    // failed) traps this kernel after ~4 s instead of hanging the GPU.
    __global__ void wait_flags_kernel(const uint32_t* flags, int count, int skip, uint32_t value)
    {
        const int i = int(threadIdx.x);
        if (i >= count || i == skip)
            return;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (true)
        {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
            if (int32_t(v - value) >= 0)
                return;
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 4000000000ull)
                __trap();
        }
    }
} // namespace glu_b200

extern "C"
{
    int glu_version(void) { return 100; }

    const char* glu_status_string(int status)
    {
        switch (status)
        {
        case GLU_SUCCESS: return "success";
        case GLU_ERROR_INVALID_ARGUMENT: return "invalid argument";
        case GLU_ERROR_INVALID_DATA_TYPE: return "invalid data type";
        case GLU_ERROR_INVALID_OPERATOR: return "invalid reduction operator";
        case GLU_ERROR_TMP_TOO_SMALL: return "temporary storage missing or too small";
        case GLU_ERROR_MISALIGNED: return "misaligned buffer";
        case GLU_ERROR_COUNT_TOO_LARGE: return "count too large";
        case GLU_ERROR_CUDA: return "CUDA error";
        default: return "unknown status";
        }
    }

    const char* glu_last_cuda_error(void) { return cudaGetErrorString(t_last_cuda_error); }

    size_t glu_data_type_size(int data_type)
    {
        DataTypeInfo info;
        if (!data_type_info(data_type, &info))
            return 0;
        return info.scalar_size * size_t(info.ncomp);
    }

    uint64_t glu_kernel_launch_count(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

    int glu_profile_enable(int on)
    {
        g_profile_on.store(on ? 1 : 0, std::memory_order_relaxed);
        return GLU_SUCCESS;
    }

    int glu_profile_collect(int kernel_id, double* total_ms, uint64_t* launches)
    {
        if (kernel_id < 0 || kernel_id >= GLU_KERNEL_COUNT_ || !total_ms || !launches)
            return GLU_ERROR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        double ms_sum = 0;
        uint64_t n = 0;
        std::vector<ProfileSpan> keep;
        for (const ProfileSpan& span : g_profile_spans)
        {
            if (span.id != kernel_id)
            {
                keep.push_back(span);
                continue;
            }
            float ms = 0;
            GLU_CUDA_TRY(cudaEventSynchronize(span.stop));
            GLU_CUDA_TRY(cudaEventElapsedTime(&ms, span.start, span.stop));
            ms_sum += ms;
            n++;
            g_profile_free_events.push_back(span.start);
            g_profile_free_events.push_back(span.stop);
        }
        g_profile_spans.swap(keep);
        *total_ms = ms_sum;
        *launches = n;
        return GLU_SUCCESS;
    }

    // ------------------------------------------------------------------------------------------ plumbing

    int glu_device_count(int* count)
    {
        if (!count)
            return GLU_ERROR_INVALID_ARGUMENT;
        *count = 0;
        GLU_CUDA_TRY(cudaGetDeviceCount(count));
        return GLU_SUCCESS;
    }

    int glu_set_device(int device)
    {
        GLU_CUDA_TRY(cudaSetDevice(device));
        return GLU_SUCCESS;
    }

    int glu_device_info(int device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor,
                        size_t* total_mem_bytes, int* warp_size)
    {
        cudaDeviceProp prop;
        GLU_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (name && name_cap)
        {
            std::strncpy(name, prop.name, name_cap - 1);
            name[name_cap - 1] = 0;
        }
        if (sm_count)
            *sm_count = prop.multiProcessorCount;
        if (cc_major)
            *cc_major = prop.major;
        if (cc_minor)
            *cc_minor = prop.minor;
        if (total_mem_bytes)
            *total_mem_bytes = prop.totalGlobalMem;
        if (warp_size)
            *warp_size = prop.warpSize;
        return GLU_SUCCESS;
    }

    int glu_malloc(void** d_ptr, size_t bytes)
    {
        if (!d_ptr)
            return GLU_ERROR_INVALID_ARGUMENT;
        *d_ptr = nullptr;
        if (bytes == 0)
            return GLU_SUCCESS;
        GLU_CUDA_TRY(cudaMalloc(d_ptr, bytes));
        return GLU_SUCCESS;
    }

    int glu_free(void* d_ptr)
    {
        if (d_ptr)
            GLU_CUDA_TRY(cudaFree(d_ptr));
        return GLU_SUCCESS;
    }

    int glu_malloc_host(void** h_ptr, size_t bytes)
    {
        if (!h_ptr)
            return GLU_ERROR_INVALID_ARGUMENT;
        *h_ptr = nullptr;
        if (bytes == 0)
            return GLU_SUCCESS;
        GLU_CUDA_TRY(cudaMallocHost(h_ptr, bytes));
        return GLU_SUCCESS;
    }

    int glu_free_host(void* h_ptr)
    {
        if (h_ptr)
            GLU_CUDA_TRY(cudaFreeHost(h_ptr));
        return GLU_SUCCESS;
    }

    int glu_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, glu_stream_t stream)
    {
        GLU_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, glu_stream_t stream)
    {
        GLU_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, glu_stream_t stream)
    {
        GLU_CUDA_TRY(
            cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_memset_u32(void* d_dst, uint32_t value, size_t count, glu_stream_t stream)
    {
        // glu/gl_utils.hpp:213-217 — ShaderStorageBuffer::clear(GLuint value)
        if (count == 0)
            return GLU_SUCCESS;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const uint32_t b = value & 0xffu;
        if (value == (b | (b << 8) | (b << 16) | (b << 24)))
            GLU_CUDA_TRY(cudaMemsetAsync(d_dst, int(b), count * sizeof(uint32_t), s));
        else
        {
            int grid = int((count + 255) / 256 < 65535 * 8 ? (count + 255) / 256 : 65535 * 8);
            fill_u32_kernel<<<grid, 256, 0, s>>>(static_cast<uint32_t*>(d_dst), value, count);
            GLU_LAUNCH_CHECK();
        }
        return GLU_SUCCESS;
    }

    int glu_signal_peers_u32(const uint64_t* h_flag_addrs, int count, uint32_t value, glu_stream_t stream)
    {
        if (!h_flag_addrs || count < 1 || count > 16)
            return GLU_ERROR_INVALID_ARGUMENT;
        glu_b200::PeerFlags f{};
        for (int i = 0; i < count; i++)
        {
            if (h_flag_addrs[i] % sizeof(uint32_t) != 0)
                return GLU_ERROR_MISALIGNED;
            f.ptr[i] = reinterpret_cast<uint32_t*>(uintptr_t(h_flag_addrs[i]));
        }
        glu_b200::signal_peers_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, count, value);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }

    int glu_stream_wait_flags_u32(const uint32_t* d_flags, int count, int skip, uint32_t value, glu_stream_t stream)
    {
        if (!d_flags || count < 1 || count > 32)
            return GLU_ERROR_INVALID_ARGUMENT;
        if (reinterpret_cast<uintptr_t>(d_flags) % sizeof(uint32_t) != 0)
            return GLU_ERROR_MISALIGNED;
        glu_b200::wait_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(d_flags, count, skip, value);
        GLU_LAUNCH_CHECK();
        return GLU_SUCCESS;
    }

    int glu_stream_create(glu_stream_t* stream)
    {
        if (!stream)
            return GLU_ERROR_INVALID_ARGUMENT;
        // a BLOCKING stream: the legacy default stream (what glu::DeviceBuffer's upload / download / clear use,
        // include/glu/device_utils.hpp) synchronises with it, so get_data() after a call enqueued on this stream
        // returns that call's result, as a GL buffer read-back after a dispatch does
        cudaStream_t s;
        GLU_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamDefault));
        *stream = s;
        return GLU_SUCCESS;
    }

    int glu_stream_destroy(glu_stream_t stream)
    {
        GLU_CUDA_TRY(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_stream_synchronize(glu_stream_t stream)
    {
        GLU_CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_event_create(glu_event_t* event)
    {
        if (!event)
            return GLU_ERROR_INVALID_ARGUMENT;
        cudaEvent_t e;
        GLU_CUDA_TRY(cudaEventCreate(&e));
        *event = e;
        return GLU_SUCCESS;
    }

    int glu_event_destroy(glu_event_t event)
    {
        GLU_CUDA_TRY(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
        return GLU_SUCCESS;
    }

    int glu_event_record(glu_event_t event, glu_stream_t stream)
    {
        GLU_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
        return GLU_SUCCESS;
    }

    int glu_event_synchronize(glu_event_t event)
    {
        GLU_CUDA_TRY(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
        return GLU_SUCCESS;
    }

    int glu_event_elapsed_ms(float* ms, glu_event_t start, glu_event_t stop)
    {
        if (!ms)
            return GLU_ERROR_INVALID_ARGUMENT;
        GLU_CUDA_TRY(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
        return GLU_SUCCESS;
    }

    // ------------------------------------------------------------------------- host-buffer entry points

    namespace
    {
        // Grow-only device buffers reused by the synchronous host-buffer entry points, one set per
        // device (the role of the ShaderStorageBuffer objects a caller of the reference keeps alive).
        struct HostPathBuffers
        {
            void* p[3] = {nullptr, nullptr, nullptr};
            size_t cap[3] = {0, 0, 0};
        };
        std::mutex g_host_path_mutex;
        HostPathBuffers g_host_path[64];

        int host_path_ensure(int slot, size_t bytes, void** out)
        {
            int dev = 0;
            GLU_CUDA_TRY(cudaGetDevice(&dev));
            if (dev < 0 || dev >= 64)
                return GLU_ERROR_INVALID_ARGUMENT;
            HostPathBuffers& b = g_host_path[dev];
            if (b.cap[slot] < bytes)
            {
                if (b.p[slot])
                    GLU_CUDA_TRY(cudaFree(b.p[slot]));
                b.p[slot] = nullptr;
                b.cap[slot] = 0;
                GLU_CUDA_TRY(cudaMalloc(&b.p[slot], bytes));
                b.cap[slot] = bytes;
            }
            *out = b.p[slot];
            return GLU_SUCCESS;
        }
#define GLU_TRY(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc__ = (expr);                                                                                             \
        if (rc__ != GLU_SUCCESS)                                                                                       \
            return rc__;                                                                                               \
    } while (0)
    } // namespace

    int glu_reduce_host(void* h_data, size_t count, int data_type, int op)
    {
        size_t esz = glu_data_type_size(data_type);
        if (esz == 0)
            return GLU_ERROR_INVALID_DATA_TYPE;
        if (!h_data || count == 0)
            return GLU_ERROR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lock(g_host_path_mutex);
        void *data = nullptr, *tmp = nullptr;
        size_t tmp_bytes = glu_reduce_tmp_bytes(count, data_type);
        GLU_TRY(host_path_ensure(0, count * esz, &data));
        GLU_TRY(host_path_ensure(2, tmp_bytes, &tmp));
        GLU_CUDA_TRY(cudaMemcpyAsync(data, h_data, count * esz, cudaMemcpyHostToDevice, 0));
        GLU_TRY(glu_reduce(data, count, data_type, op, tmp, tmp_bytes, nullptr));
        GLU_CUDA_TRY(cudaMemcpyAsync(h_data, data, esz, cudaMemcpyDeviceToHost, 0));
        GLU_CUDA_TRY(cudaStreamSynchronize(0));
        return GLU_SUCCESS;
    }

    int glu_scan_exclusive_host(void* h_data, size_t count, size_t num_partitions, int data_type)
    {
        size_t esz = glu_data_type_size(data_type);
        if (esz == 0)
            return GLU_ERROR_INVALID_DATA_TYPE;
        if (!h_data || count == 0 || num_partitions == 0)
            return GLU_ERROR_INVALID_ARGUMENT;
        std::lock_guard<std::mutex> lock(g_host_path_mutex);
        void *data = nullptr, *tmp = nullptr;
        size_t bytes = count * num_partitions * esz;
        size_t tmp_bytes = glu_scan_exclusive_tmp_bytes(count, num_partitions, data_type);
        GLU_TRY(host_path_ensure(0, bytes, &data));
        GLU_TRY(host_path_ensure(2, tmp_bytes, &tmp));
        GLU_CUDA_TRY(cudaMemcpyAsync(data, h_data, bytes, cudaMemcpyHostToDevice, 0));
        GLU_TRY(glu_scan_exclusive(data, count, num_partitions, data_type, tmp, tmp_bytes, nullptr));
        GLU_CUDA_TRY(cudaMemcpyAsync(h_data, data, bytes, cudaMemcpyDeviceToHost, 0));
        GLU_CUDA_TRY(cudaStreamSynchronize(0));
        return GLU_SUCCESS;
    }

    int glu_radix_sort_u32kv_host(uint32_t* h_keys, uint32_t* h_vals, size_t count, size_t num_steps)
    {
        if (!h_keys || !h_vals)
            return GLU_ERROR_INVALID_ARGUMENT;
        if (count <= 1)
            return GLU_SUCCESS;
        size_t bytes = count * sizeof(uint32_t);
        size_t tmp_bytes = glu_radix_sort_u32kv_tmp_bytes(count);
        if (tmp_bytes == 0)
            return GLU_ERROR_COUNT_TOO_LARGE;
        std::lock_guard<std::mutex> lock(g_host_path_mutex);
        void *keys = nullptr, *vals = nullptr, *tmp = nullptr;
        GLU_TRY(host_path_ensure(0, bytes, &keys));
        GLU_TRY(host_path_ensure(1, bytes, &vals));
        GLU_TRY(host_path_ensure(2, tmp_bytes, &tmp));
        GLU_CUDA_TRY(cudaMemcpyAsync(keys, h_keys, bytes, cudaMemcpyHostToDevice, 0));
        GLU_CUDA_TRY(cudaMemcpyAsync(vals, h_vals, bytes, cudaMemcpyHostToDevice, 0));
        GLU_TRY(glu_radix_sort_u32kv(static_cast<uint32_t*>(keys), static_cast<uint32_t*>(vals), count, num_steps, tmp,
                                     tmp_bytes, nullptr));
        GLU_CUDA_TRY(cudaMemcpyAsync(h_keys, keys, bytes, cudaMemcpyDeviceToHost, 0));
        GLU_CUDA_TRY(cudaMemcpyAsync(h_vals, vals, bytes, cudaMemcpyDeviceToHost, 0));
        GLU_CUDA_TRY(cudaStreamSynchronize(0));
        return GLU_SUCCESS;
    }
}

// ---- queue of host-buffer sorts (upload / sort / download of consecutive jobs overlap) ---------------------------
struct glu_host_sort_queue
{
    struct Slot
    {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        uint32_t* keys = nullptr;
        uint32_t* vals = nullptr;
        void* tmp = nullptr;
        bool busy = false;
    };
    int device = 0;
    size_t max_count = 0, tmp_bytes = 0;
    std::vector<Slot> slots;
    size_t next = 0;
};

extern "C" int glu_host_sort_queue_destroy(glu_host_sort_queue_t* q)
{
    if (!q)
        return GLU_ERROR_INVALID_ARGUMENT;
    int rc = GLU_SUCCESS;
    for (auto& s : q->slots)
    {
        if (s.stream && cudaStreamSynchronize(s.stream) != cudaSuccess)
            rc = GLU_ERROR_CUDA;
        if (s.keys)
            cudaFree(s.keys);
        if (s.vals)
            cudaFree(s.vals);
        if (s.tmp)
            cudaFree(s.tmp);
        if (s.done)
            cudaEventDestroy(s.done);
        if (s.stream)
            cudaStreamDestroy(s.stream);
    }
    delete q;
    return rc;
}

extern "C" int glu_host_sort_queue_create(glu_host_sort_queue_t** out, size_t max_count, int depth)
{
    if (!out || depth < 1 || depth > 8 || max_count == 0)
        return GLU_ERROR_INVALID_ARGUMENT;
    const size_t tmp_bytes = glu_radix_sort_u32kv_tmp_bytes(max_count);
    if (tmp_bytes == 0)
        return GLU_ERROR_COUNT_TOO_LARGE;
    glu_host_sort_queue* q = new (std::nothrow) glu_host_sort_queue;
    if (!q)
        return GLU_ERROR_CUDA;
    q->max_count = max_count;
    q->tmp_bytes = tmp_bytes;
    q->slots.resize(size_t(depth));
    bool ok = cudaGetDevice(&q->device) == cudaSuccess;
    for (auto& s : q->slots)
    {
        ok = ok && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(reinterpret_cast<void**>(&s.keys), max_count * sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaMalloc(reinterpret_cast<void**>(&s.vals), max_count * sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaMalloc(&s.tmp, tmp_bytes) == cudaSuccess;
    }
    if (!ok)
    {
        cudaGetLastError();
        glu_host_sort_queue_destroy(q);
        return GLU_ERROR_CUDA;
    }
    *out = q;
    return GLU_SUCCESS;
}

extern "C" int glu_host_sort_queue_submit(glu_host_sort_queue_t* q, uint32_t* h_keys, uint32_t* h_vals, size_t count,
                                          size_t num_steps)
{
    if (!q || !h_keys || !h_vals)
        return GLU_ERROR_INVALID_ARGUMENT;
    if (count > q->max_count)
        return GLU_ERROR_COUNT_TOO_LARGE;
    if (count <= 1) // glu/RadixSort.hpp:278-279
        return GLU_SUCCESS;
    glu_host_sort_queue::Slot& s = q->slots[q->next];
    q->next = (q->next + 1) % q->slots.size();
    if (s.busy)
        GLU_CUDA_TRY(cudaEventSynchronize(s.done));
    s.busy = false;
    const size_t bytes = count * sizeof(uint32_t);
    GLU_CUDA_TRY(cudaMemcpyAsync(s.keys, h_keys, bytes, cudaMemcpyHostToDevice, s.stream));
    GLU_CUDA_TRY(cudaMemcpyAsync(s.vals, h_vals, bytes, cudaMemcpyHostToDevice, s.stream));
    const int rc = glu_radix_sort_u32kv(s.keys, s.vals, count, num_steps, s.tmp, q->tmp_bytes, s.stream);
    if (rc != GLU_SUCCESS)
        return rc;
    GLU_CUDA_TRY(cudaMemcpyAsync(h_keys, s.keys, bytes, cudaMemcpyDeviceToHost, s.stream));
    GLU_CUDA_TRY(cudaMemcpyAsync(h_vals, s.vals, bytes, cudaMemcpyDeviceToHost, s.stream));
    GLU_CUDA_TRY(cudaEventRecord(s.done, s.stream));
    s.busy = true;
    return GLU_SUCCESS;
}

extern "C" int glu_host_sort_queue_wait(glu_host_sort_queue_t* q)
{
    if (!q)
        return GLU_ERROR_INVALID_ARGUMENT;
    for (auto& s : q->slots)
    {
        if (s.busy)
            GLU_CUDA_TRY(cudaEventSynchronize(s.done));
        s.busy = false;
    }
    return GLU_SUCCESS;
}

// ---- CUDA IPC (one process per GPU: map a peer rank's receive buffer into this process) ----------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == GLU_IPC_HANDLE_BYTES, "handle size");

extern "C" int glu_ipc_get_handle(void* d_ptr, unsigned char handle[GLU_IPC_HANDLE_BYTES])
{
    if (!d_ptr || !handle)
        return GLU_ERROR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    GLU_CUDA_TRY(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, sizeof h);
    return GLU_SUCCESS;
}

extern "C" int glu_ipc_open_handle(const unsigned char handle[GLU_IPC_HANDLE_BYTES], void** d_ptr)
{
    if (!handle || !d_ptr)
        return GLU_ERROR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    GLU_CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GLU_SUCCESS;
}

extern "C" int glu_ipc_close_handle(void* d_ptr)
{
    if (!d_ptr)
        return GLU_ERROR_INVALID_ARGUMENT;
    GLU_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return GLU_SUCCESS;
}
