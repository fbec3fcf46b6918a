// cub_compare.cu — same-box comparator, NOT part of the product and not linked into it: times CUB's
// DeviceRadixSort::SortPairs / DeviceScan::ExclusiveSum / DeviceReduce::Sum (the CUDA toolkit's own library,
// /usr/local/cuda/include/cub) on the workloads bench.py measures, so that the library's numbers can be read
// next to the best generally available implementation on the same GPU.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o cub_compare cub_compare.cu
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                          \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e = (x);                                                                                           \
        if (e != cudaSuccess)                                                                                          \
        {                                                                                                              \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                                               \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

__global__ void fill(uint32_t* keys, uint32_t* vals, size_t n, uint32_t seed)
{
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        uint32_t x = uint32_t(i) * 2654435761u + seed; // counter-based hash (murmur3 finaliser)
        x ^= x >> 16;
        x *= 0x85ebca6bu;
        x ^= x >> 13;
        x *= 0xc2b2ae35u;
        x ^= x >> 16;
        keys[i] = x;
        vals[i] = uint32_t(i);
    }
}

template<typename F> static float median_ms(F&& run, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    std::vector<float> t;
    for (int i = 0; i < reps + 3; i++)
    {
        float ms = run(a, b);
        if (i >= 3)
            t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv)
{
    const int log2n = argc > 1 ? std::atoi(argv[1]) : 28;
    const size_t n = size_t(1) << log2n;
    uint32_t *k0, *v0, *k1, *v1, *k2, *v2;
    CK(cudaMalloc(&k0, 4 * n));
    CK(cudaMalloc(&v0, 4 * n));
    CK(cudaMalloc(&k1, 4 * n));
    CK(cudaMalloc(&v1, 4 * n));
    CK(cudaMalloc(&k2, 4 * n));
    CK(cudaMalloc(&v2, 4 * n));
    fill<<<1184, 512>>>(k0, v0, n, 1u);
    CK(cudaDeviceSynchronize());

    // ---- SortPairs (DoubleBuffer: the same ping-pong budget as glu_radix_sort_u32kv)
    {
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<uint32_t> dk(k1, k2), dv(v1, v2);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n));
        void* tmp;
        CK(cudaMalloc(&tmp, tmp_bytes));
        float ms = median_ms(
            [&](cudaEvent_t a, cudaEvent_t b) {
                CK(cudaMemcpy(k1, k0, 4 * n, cudaMemcpyDeviceToDevice));
                CK(cudaMemcpy(v1, v0, 4 * n, cudaMemcpyDeviceToDevice));
                cub::DoubleBuffer<uint32_t> kk(k1, k2), vv(v1, v2);
                CK(cudaEventRecord(a));
                CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kk, vv, n));
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float t;
                CK(cudaEventElapsedTime(&t, a, b));
                return t;
            },
            10);
        std::printf("cub::DeviceRadixSort::SortPairs  u32/u32 n=2^%d: %.3f ms  %.2f Gpairs/s\n", log2n, ms, n / ms / 1e6);
        CK(cudaFree(tmp));
    }
    // ---- ExclusiveSum
    {
        size_t tmp_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, k1, k2, n));
        void* tmp;
        CK(cudaMalloc(&tmp, tmp_bytes));
        float ms = median_ms(
            [&](cudaEvent_t a, cudaEvent_t b) {
                CK(cudaMemcpy(k1, k0, 4 * n, cudaMemcpyDeviceToDevice));
                CK(cudaEventRecord(a));
                CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, k1, k1, n)); // in place, like glu_scan_exclusive
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float t;
                CK(cudaEventElapsedTime(&t, a, b));
                return t;
            },
            10);
        std::printf("cub::DeviceScan::ExclusiveSum    u32 n=2^%d: %.3f ms  %.0f GB/s (8 B/elem)\n", log2n, ms, 8.0 * n / ms / 1e6);
        CK(cudaFree(tmp));
    }
    // ---- Sum
    {
        size_t tmp_bytes = 0;
        CK(cub::DeviceReduce::Sum(nullptr, tmp_bytes, k1, k2, n));
        void* tmp;
        CK(cudaMalloc(&tmp, tmp_bytes));
        float ms = median_ms(
            [&](cudaEvent_t a, cudaEvent_t b) {
                CK(cudaMemcpy(k1, k0, 4 * n, cudaMemcpyDeviceToDevice));
                CK(cudaEventRecord(a));
                CK(cub::DeviceReduce::Sum(tmp, tmp_bytes, k1, k2, n));
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float t;
                CK(cudaEventElapsedTime(&t, a, b));
                return t;
            },
            10);
        std::printf("cub::DeviceReduce::Sum           u32 n=2^%d: %.3f ms  %.0f GB/s (4 B/elem)\n", log2n, ms, 4.0 * n / ms / 1e6);
        CK(cudaFree(tmp));
    }
    return 0;
}
