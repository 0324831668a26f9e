// flagstat_kernel_tma.cuh -- the flagstat / pospopcnt kernel with the input
// staged through shared memory by the TMA unit (cp.async.bulk + mbarrier).
//
// Same arithmetic as flagstat_kernels.cuh (mask select + bit-sliced counters);
// what changes is how bytes get from HBM to the registers:
//
//   producer  one elected thread of an extra warp issues 16 KiB bulk copies
//             global -> shared into a ring of kStages buffers, each guarded by a
//             "full" mbarrier (armed with the byte count) and an "empty"
//             mbarrier (one arrival per consumer warp);
//   consumers the 8 compute warps wait on "full", pull their 4 x 16 bytes with
//             conflict-free LDS.128, release the slot, and run the batch step.
//
// The loads in flight now live in shared memory instead of registers: the
// consumer needs no double buffer (32 registers less), more CTAs fit per SM, and
// kStages x 16 KiB per CTA are outstanding regardless of how busy the integer
// pipe keeps the warps -- which is what a kernel that sits at ~80 % ALU-pipe
// utilisation needs to stop stalling on long-scoreboard waits (see
// profiles/r1a_ncu_summary.md).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "flagstat_kernels.cuh"

namespace fsb200 {

constexpr int kStageBytes = kVecPerBatch * 16;  // 16 KiB: one CTA batch

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <int MODE, int VARIANT, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads + 32, MINB)
flagstat_kernel_tma(const uint16_t* __restrict__ base, uint64_t n,
                    unsigned long long* __restrict__ out, const __grid_constant__ XchgArgs xa)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* ring = reinterpret_cast<uint4*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * kStageBytes);
    uint64_t* empty = full + STAGES;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t addr = reinterpret_cast<uint64_t>(base);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint4* __restrict__ body = reinterpret_cast<const uint4*>(base + head);
    const uint64_t V = (n - head) >> 3;
    const uint64_t tail_start = head + (V << 3);
    const uint64_t ntail = n - tail_start;
    const uint64_t NB = V / kVecPerBatch;
    const uint64_t G = gridDim.x;
    const uint32_t my = (NB > blockIdx.x) ? (uint32_t)((NB - blockIdx.x + G - 1) / G) : 0u;
    const uint64_t stride = G * (uint64_t)kVecPerBatch;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    unsigned long long acc_all = 0ull, acc_fail = 0ull;

    if (warp == kWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            const uint4* src = body + (uint64_t)blockIdx.x * kVecPerBatch;
            for (uint32_t it = 0; it < my; ++it) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], kStageBytes);
                bulk_g2s(ring + (size_t)s * kVecPerBatch, src, kStageBytes, &full[s]);
                src += stride;
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        Lanes<MODE, VARIANT> st;
        st.clear();
        uint32_t b = 0;

        if (blockIdx.x == (uint32_t)(NB % G)) {
            uint32_t w[16];
            {
                uint4 v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const uint64_t idx = NB * kVecPerBatch + (uint64_t)u * kThreads + tid;
                    v[u] = (idx < V) ? ld_stream(body + idx) : make_uint4(0u, 0u, 0u, 0u);
                }
                unpack4(v, w);
            }
            st.step(w, b);
            ++b;
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = 0u;
            if (tid < head) w[0] = base[tid];
            else if (tid - head < ntail) w[0] = base[tail_start + (tid - head)];
            st.step(w, b);
            ++b;
        }

        uint32_t it = 0, s = 0, ph = 0;
        do {
            while (it < my && b < Counter::kMaxBatches) {
                mbar_wait(&full[s], ph);
                uint32_t w[16];
                {
                    const uint4* src = ring + (size_t)s * kVecPerBatch + tid;
                    uint4 v[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u) v[u] = src[u * kThreads];
                    unpack4(v, w);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
                st.step(w, b);
                ++b;
                ++it;
            }
            acc_all += st.all.flush_warp(b, lane);
            if (MODE == kFlagstat && st.nfail != 0u)
                acc_fail += st.fail.flush_warp(st.nfail, lane);
            st.clear();
            b = 0;
        } while (it < my);
    }

    cta_epilogue<MODE>(out, acc_all, acc_fail, n, warp, lane, warp < kWarps, xa);
}


// ---------------------------------------------------------------------------
// Thread-private cp.async ring (variants 6, 7)
//
// Every thread prefetches ITS OWN 4 x 16 bytes of the next DEPTH batches into a
// private slice of shared memory with cp.async (SASS LDGSTS) and reads them back
// with LDS.128 once cp.async.wait_group says its own copies have landed.  No
// mbarrier, no producer warp, no CTA-wide synchronisation in the loop: a thread
// only ever consumes what it fetched itself.  The bytes in flight live in
// shared memory (DEPTH x 16 KiB per CTA), not in registers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src)
{
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst_smem), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int MODE, int VARIANT, int DEPTH, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
flagstat_kernel_ring(const uint16_t* __restrict__ base, uint64_t n,
                     unsigned long long* __restrict__ out, const __grid_constant__ XchgArgs xa)
{
    static_assert((DEPTH & (DEPTH - 1)) == 0, "DEPTH must be a power of two");
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t addr = reinterpret_cast<uint64_t>(base);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint4* __restrict__ body = reinterpret_cast<const uint4*>(base + head);
    const uint64_t V = (n - head) >> 3;
    const uint64_t tail_start = head + (V << 3);
    const uint64_t ntail = n - tail_start;
    const uint64_t NB = V / kVecPerBatch;
    const uint64_t G = gridDim.x;
    const uint32_t my = (NB > blockIdx.x) ? (uint32_t)((NB - blockIdx.x + G - 1) / G) : 0u;
    const uint64_t stride = G * (uint64_t)kVecPerBatch;

    Lanes<MODE, VARIANT> st;
    st.clear();
    uint32_t b = 0;
    unsigned long long acc_all = 0ull, acc_fail = 0ull;

    if (blockIdx.x == (uint32_t)(NB % G)) {
        uint32_t w[16];
        {
            uint4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint64_t idx = NB * kVecPerBatch + (uint64_t)u * kThreads + tid;
                v[u] = (idx < V) ? ld_stream(body + idx) : make_uint4(0u, 0u, 0u, 0u);
            }
            unpack4(v, w);
        }
        st.step(w, b);
        ++b;
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
        if (tid < head) w[0] = base[tid];
        else if (tid - head < ntail) w[0] = base[tail_start + (tid - head)];
        st.step(w, b);
        ++b;
    }

    // this thread's slice of stage s, load u:  ((s * kU + u) * kThreads + tid) * 16 bytes
    const uint32_t my_smem = smem_u32(smem_raw) + tid * 16u;
    const uint4* __restrict__ src = body + tid + (uint64_t)blockIdx.x * kVecPerBatch;

    // prologue: DEPTH batches in flight (empty groups keep the group count uniform)
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
        if ((uint32_t)d < my) {
#pragma unroll
            for (int u = 0; u < kU; ++u)
                cp_async16(my_smem + (uint32_t)((d * kU + u) * kThreads * 16), src + u * kThreads);
            src += stride;
        }
        cp_async_commit();
    }

    uint32_t it = 0;
    do {
        while (it < my && b < Counter::kMaxBatches) {
            cp_async_wait<DEPTH - 1>();
            const uint32_t stage = my_smem + (it & (uint32_t)(DEPTH - 1)) * (uint32_t)kStageBytes;
            uint32_t w[16];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(w[4 * u]), "=r"(w[4 * u + 1]), "=r"(w[4 * u + 2]),
                               "=r"(w[4 * u + 3])
                             : "r"(stage + (uint32_t)(u * kThreads * 16)));
            st.step(w, b);
            // the registers have been consumed: this slot can take batch it + DEPTH
            if (it + (uint32_t)DEPTH < my) {
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    cp_async16(stage + (uint32_t)(u * kThreads * 16), src + u * kThreads);
                src += stride;
            }
            cp_async_commit();
            ++b;
            ++it;
        }
        acc_all += st.all.flush_warp(b, lane);
        if (MODE == kFlagstat && st.nfail != 0u) acc_fail += st.fail.flush_warp(st.nfail, lane);
        st.clear();
        b = 0;
    } while (it < my);
    cp_async_wait<0>();

    cta_epilogue<MODE>(out, acc_all, acc_fail, n, warp, lane, true, xa);
}

}  // namespace fsb200
