// hbm_read_probe.cu -- what a READ-ONLY stream reaches on this B200, next to a copy.
//
// MEASURED_PEAKS.json (the roofline denominator bench.py reports against) is a COPY
// kernel: read + write bytes.  The flagstat kernel only reads, and a read-only stream
// does not pay the bus turn-arounds of a copy, so its fraction of that peak can exceed 1.
// This probe measures the read-only ceiling itself with kernels that do (almost) no
// arithmetic -- one XOR per 16 bytes -- over the same 1,649,083,784 bytes as the
// headline workload:
//   ldg      ld.global.nc.L1::no_allocate.v4, U loads in flight per thread
//   ring     the flagstat kernel's own load path: thread-private cp.async ring, depth 4,
//            4 x 16 B per stage, LDS.128 back (flagstat_kernel_group.cuh)
//   copy     uint4 load + store (bytes = 2 x size), the MEASURED_PEAKS.json shape
// Prints GB/s (median and best of 20 launches, CUDA events).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/hbm_read_probe tools/hbm_read_probe.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            std::exit(1);                                                             \
        }                                                                             \
    } while (0)

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

template <int U>
__global__ void __launch_bounds__(kThreads) k_ldg(const uint4* __restrict__ p, uint64_t nvec,
                                                  uint32_t* __restrict__ sink)
{
    uint32_t acc = 0;
    const uint64_t per = (uint64_t)U * kThreads;
    const uint64_t nb = nvec / per;
    for (uint64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(p + b * per + (uint64_t)u * kThreads + threadIdx.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

template <int DEPTH, int U>
__global__ void __launch_bounds__(kThreads) k_ring(const uint4* __restrict__ p, uint64_t nvec,
                                                   uint32_t* __restrict__ sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t my = smem_u32(smem) + threadIdx.x * 16u;
    const uint64_t per = (uint64_t)U * kThreads;
    const uint64_t nb = nvec / per;
    const uint64_t G = gridDim.x;
    const uint32_t mine = nb > blockIdx.x ? (uint32_t)((nb - blockIdx.x + G - 1) / G) : 0u;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p + blockIdx.x * per + threadIdx.x);
    const uint64_t stride = G * per * 16u;
    uint32_t fetched = 0, acc = 0;
    auto fetch = [&](int s) {
        if (fetched < mine) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(
                                 my + (uint32_t)((s * U + u) * kThreads * 16)),
                             "l"(src + (uint64_t)u * kThreads * 16)
                             : "memory");
            src += stride;
            ++fetched;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < DEPTH; ++s) fetch(s);
    for (uint32_t b = 0; b < mine; b += DEPTH) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) {
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            if (b + s < mine) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t a, bb, c, d;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(a), "=r"(bb), "=r"(c), "=r"(d)
                                 : "r"(my + (uint32_t)((s * U + u) * kThreads * 16)));
                    acc ^= a ^ bb ^ c ^ d;
                }
            }
            fetch(s);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void __launch_bounds__(kThreads) k_copy(const uint4* __restrict__ p, uint4* __restrict__ q,
                                                   uint64_t nvec)
{
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec;
         i += (uint64_t)gridDim.x * kThreads)
        q[i] = p[i];
}

template <class F>
static void timeit(const char* name, double bytes, F&& launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i) launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ms;
    for (int i = 0; i < 20; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        ms.push_back(t);
    }
    CK(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    std::printf("%-44s median %8.1f GB/s   best %8.1f GB/s   (%.1f us)\n", name,
                bytes / (ms[ms.size() / 2] * 1e6), bytes / (ms[0] * 1e6), ms[ms.size() / 2] * 1e3);
}

int main()
{
    const uint64_t bytes = 1649083784ull & ~15ull;
    const uint64_t nvec = bytes / 16;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    std::printf("device: %s, %d SMs, %.3f GB per pass\n", pr.name, sms, bytes / 1e9);
    uint4 *a, *b;
    uint32_t* sink;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 0x5a, bytes));
    CK(cudaMemset(b, 0, bytes));
    char name[96];
    for (int per_sm : {1, 2, 4, 8}) {
        std::snprintf(name, sizeof name, "read ldg U=4, %d CTA/SM", per_sm);
        timeit(name, (double)bytes, [&] { k_ldg<4><<<sms * per_sm, kThreads>>>(a, nvec, sink); });
        std::snprintf(name, sizeof name, "read ldg U=8, %d CTA/SM", per_sm);
        timeit(name, (double)bytes, [&] { k_ldg<8><<<sms * per_sm, kThreads>>>(a, nvec, sink); });
    }
    CK(cudaFuncSetAttribute(k_ring<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * kThreads * 16));
    CK(cudaFuncSetAttribute(k_ring<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 4 * kThreads * 16));
    CK(cudaFuncSetAttribute(k_ring<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8 * kThreads * 16));
    for (int per_sm : {1, 2, 3}) {
        std::snprintf(name, sizeof name, "read cp.async ring depth 4 x 64 B, %d CTA/SM", per_sm);
        timeit(name, (double)bytes,
               [&] { k_ring<4, 4><<<sms * per_sm, kThreads, 4 * 4 * kThreads * 16>>>(a, nvec, sink); });
    }
    timeit("read cp.async ring depth 6 x 64 B, 2 CTA/SM", (double)bytes,
           [&] { k_ring<6, 4><<<sms * 2, kThreads, 6 * 4 * kThreads * 16>>>(a, nvec, sink); });
    timeit("read cp.async ring depth 4 x 128 B, 1 CTA/SM", (double)bytes,
           [&] { k_ring<4, 8><<<sms * 1, kThreads, 4 * 8 * kThreads * 16>>>(a, nvec, sink); });
    for (int per_sm : {4, 8, 16})  {
        std::snprintf(name, sizeof name, "copy uint4 (read+write bytes), %d CTA/SM", per_sm);
        timeit(name, 2.0 * bytes, [&] { k_copy<<<sms * per_sm, kThreads>>>(a, b, nvec); });
    }
    timeit("cudaMemcpyAsync D2D (read+write bytes)", 2.0 * bytes,
           [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, 0); });
    return 0;
}
