// flagstat_kernel_group.cuh -- the default flagstat / pospopcnt kernel: the
// thread-private cp.async ring of flagstat_kernel_tma.cuh with the counters
// restructured so that NOTHING in the hot loop depends on a run-time batch index.
//
// The first ring kernel carried its weight-16 words through "hold" levels chosen
// by the bits of the batch counter.  ncu / SASS showed what that costs: per
// 16-word batch and counter ~27 predicated moves, PLOP3s and branches on top of
// the 30 LOP3 of the adder tree (profiles/r1n_*).  Here four batches -- exactly
// one lap of the depth-4 ring, so shared-memory stage offsets are immediates
// too -- form a GROUP that is unrolled at compile time:
//
//   POS 0  tree(16 words) -> c      h16 = c
//   POS 1  tree -> c                (c32, p16) = p16 + h16 + c
//   POS 2  tree -> c                h16 = c
//   POS 3  tree -> c                (t, p16) = p16 + h16 + c
//                                   (c64, p32) = p32 + c32 + t ;  ripple c64 into up[]
//
// 4 x 15 + 3 full adders and a 6-plane half-adder ripple per 64 words = 2.2 LOP3
// per word, and that is now also what the generated code executes.
//
// The QC-fail counter sees a data-dependent subsequence of the batches, so it
// follows the same positions: a batch without QC-fail records either skips it
// entirely (nothing pending: the only case in real data) or feeds it zeros
// (absorb_zero, 0-4 LOP3).
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "flagstat_kernel_tma.cuh"

namespace fsb200 {

template <int NUP>
struct GroupCounter {
    static constexpr int kPlanes = 6 + NUP;
    // a position can hold 2^kPlanes - 1; one group adds at most 64
    static constexpr uint32_t kMaxGroups = ((1u << kPlanes) - 1u) / 64u;

    uint32_t p1, p2, p4, p8, p16, p32;
    uint32_t h16, c32;  // pending words of weight 16 / 32; zero unless pending
    uint32_t up[NUP];

    __device__ __forceinline__ void clear()
    {
        p1 = p2 = p4 = p8 = p16 = p32 = h16 = c32 = 0u;
#pragma unroll
        for (int k = 0; k < NUP; ++k) up[k] = 0u;
    }

    // Harley-Seal over 16 words into p1,p2,p4,p8; returns the carry of weight 16
    __device__ __forceinline__ uint32_t tree(const uint32_t (&x)[16])
    {
        uint32_t t2a, t2b, t4a, t4b, t8a, t8b, c;
        FSB_CSA(t2a, p1, p1, x[0], x[1]);
        FSB_CSA(t2b, p1, p1, x[2], x[3]);
        FSB_CSA(t4a, p2, p2, t2a, t2b);
        FSB_CSA(t2a, p1, p1, x[4], x[5]);
        FSB_CSA(t2b, p1, p1, x[6], x[7]);
        FSB_CSA(t4b, p2, p2, t2a, t2b);
        FSB_CSA(t8a, p4, p4, t4a, t4b);
        FSB_CSA(t2a, p1, p1, x[8], x[9]);
        FSB_CSA(t2b, p1, p1, x[10], x[11]);
        FSB_CSA(t4a, p2, p2, t2a, t2b);
        FSB_CSA(t2a, p1, p1, x[12], x[13]);
        FSB_CSA(t2b, p1, p1, x[14], x[15]);
        FSB_CSA(t4b, p2, p2, t2a, t2b);
        FSB_CSA(t8b, p4, p4, t4a, t4b);
        FSB_CSA(c, p8, p8, t8a, t8b);
        return c;
    }

    __device__ __forceinline__ void ripple(uint32_t c)  // c has weight 64
    {
#pragma unroll
        for (int k = 0; k < NUP; ++k) {
            const uint32_t t = up[k] & c;
            up[k] ^= c;
            c = t;
        }
    }

    // what a batch does with its weight-16 carry c, by position in the group
    template <int POS>
    __device__ __forceinline__ void place(uint32_t c)
    {
        if (POS == 0 || POS == 2) {
            h16 = c;
        } else if (POS == 1) {
            FSB_CSA(c32, p16, p16, h16, c);
            h16 = 0u;
        } else {
            uint32_t t, c64;
            FSB_CSA(t, p16, p16, h16, c);
            FSB_CSA(c64, p32, p32, c32, t);
            ripple(c64);
            h16 = 0u;
            c32 = 0u;
        }
    }

    template <int POS>
    __device__ __forceinline__ void absorb(const uint32_t (&x)[16])
    {
        place<POS>(tree(x));
    }

    // a batch of sixteen zero words
    template <int POS>
    __device__ __forceinline__ void absorb_zero()
    {
        if (POS == 0 || POS == 2) {
            h16 = 0u;
        } else if (POS == 1) {
            c32 = p16 & h16;
            p16 ^= h16;
            h16 = 0u;
        } else {
            close();
        }
    }

    // end a (possibly partial) group: fold the pending words into the planes
    __device__ __forceinline__ void close()
    {
        const uint32_t t = p16 & h16;
        p16 ^= h16;
        uint32_t c64;
        FSB_CSA(c64, p32, p32, c32, t);
        ripple(c64);
        h16 = 0u;
        c32 = 0u;
    }

    // Warp-cooperative expansion of a CLOSED counter: lane j gets the warp-wide
    // total of bit position j (same butterfly as BitCounter::flush_warp).
    __device__ __forceinline__ uint32_t flush_warp(uint32_t lane) const
    {
        uint32_t N[kPlanes + 5];
        N[0] = p1; N[1] = p2; N[2] = p4; N[3] = p8; N[4] = p16; N[5] = p32;
#pragma unroll
        for (int k = 0; k < NUP; ++k) N[6 + k] = up[k];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            uint32_t carry = 0u;
#pragma unroll
            for (int k = 0; k < kPlanes + s; ++k) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, N[k], 1 << s);
                const uint32_t sum = xor3(N[k], o, carry);
                carry = maj3(N[k], o, carry);
                N[k] = sum;
            }
            N[kPlanes + s] = carry;
        }
        uint32_t v = 0u;
#pragma unroll
        for (int k = 0; k < kPlanes + 5; ++k) v |= ((N[k] >> lane) & 1u) << k;
        return v;
    }
};

using GCounter = GroupCounter<6>;  // 12 planes: 63 groups = 4032 words per epoch
constexpr uint32_t kDenseGroups = 7u;  // detect-free groups after a group that was all general-path

template <int MODE, int VARIANT>
struct GroupLanes {
    GCounter all;
    GCounter fail;    // unused in pospopcnt mode
    bool fail_open;   // `fail` holds pending words inside the current group (warp-uniform)
    bool fail_dirty;  // `fail` absorbed something this epoch (warp-uniform)

    __device__ __forceinline__ void clear()
    {
        all.clear();
        if (MODE != kPospopcnt) fail.clear();
        fail_open = false;
        fail_dirty = false;
    }

    // One batch.  `dense` (warp-uniform, VARIANT 3 only): skip the OR-detect and feed
    // the second counter unconditionally; that is correct for any data, the detect
    // only buys the cheaper path.  Returns whether the batch needed the second
    // counter (feeds the caller's dense-mode decision).
    template <int POS>
    __device__ __forceinline__ bool step(const uint32_t (&w)[16], bool dense = false)
    {
        if (MODE == kPospopcnt) {
            all.template absorb<POS>(w);
            return false;
        }
        // warp-uniform view of the batch: OR of all 512 packed words (REDUX.OR)
        uint32_t wany = 0xffffffffu;
        if (VARIANT != 3 || !dense) {
            uint32_t any = w[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) any |= w[i];
            wany = __reduce_or_sync(0xffffffffu, any);
        }
        const bool has_sec = (wany & 0x01000100u) != 0u;   // a SECONDARY record in this warp batch
        const bool has_fail = (wany & 0x02000200u) != 0u;  // a QC-fail record

        uint32_t y[16];
        if (MODE == kSamtools) {
            // position 4 must be the clean class-K indicator on every path
            static_assert(MODE != kSamtools || VARIANT == 3, "kSamtools builds on the FMA-pipe forms");
            if (has_sec) {
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = mask_select_fx<true>(w[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) y[i] = mask_select_fx<false>(w[i]);
            }
        } else if (VARIANT == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = mask_select_i(w[i]);
        } else if (VARIANT == 3 && (has_sec || has_fail)) {
            // ALU-bound batches: keep-mask and SECONDARY fix-up built on the FMA pipe (4 ALU +
            // 4 HFMA2 per word).  QC-clean batches without SECONDARY records -- real data, where
            // DRAM is the limit -- keep the form with the fewest instructions overall (below).
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = mask_select_f(w[i]);
        } else if (!has_sec) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = mask_select_h_nosec(w[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = mask_select_h(w[i]);
        }
        all.template absorb<POS>(y);
        if (has_fail) {  // second counter
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = VARIANT == 3 ? fail_gate_f(w[i], y[i]) : (y[i] & fail_mask<VARIANT>(w[i]));
            fail.template absorb<POS>(y);
            fail_dirty = true;
            fail_open = POS != 3;
        } else if (fail_open) {
            fail.template absorb_zero<POS>();
            fail_open = POS != 3;
        }
        return VARIANT == 3 ? has_fail : (has_sec && has_fail);
    }

    __device__ __forceinline__ void close()
    {
        all.close();
        if (MODE != kPospopcnt && fail_open) {
            fail.close();
            fail_open = false;
        }
    }
};

// cp.async with the per-load offsets as immediates on both sides
template <int OFF_S, int OFF_G>
__device__ __forceinline__ void cp_async16_imm(uint32_t dst_smem, const void* src)
{
    asm volatile("cp.async.cg.shared.global.L2::128B [%0 + %2], [%1 + %3], 16;" ::"r"(dst_smem),
                 "l"(src), "n"(OFF_S), "n"(OFF_G)
                 : "memory");
}

template <int OFF>
__device__ __forceinline__ void lds128_imm(uint32_t smem, uint32_t& a, uint32_t& b, uint32_t& c,
                                           uint32_t& d)
{
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4 + %5];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(smem), "n"(OFF));
}

// Timeline probe (tools/timeline_probe.py builds a second .so with -DFSB_TIMELINE; never in the
// product build): thread 0 of every CTA stamps %globaltimer at five points of its life.
#ifdef FSB_TIMELINE
__device__ unsigned long long g_timeline[8 * 2048];
#define FSB_TL(k)                                                                      \
    do {                                                                               \
        if (threadIdx.x == 0 && blockIdx.x < 2048u) {                                  \
            g_timeline[blockIdx.x * 8u + (k)] = global_timer_ns();                     \
            if ((k) == 0) {                                                            \
                uint32_t sm_;                                                          \
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));                       \
                g_timeline[blockIdx.x * 8u + 6u] = sm_ + 1u;                           \
            }                                                                          \
        }                                                                              \
    } while (0)
#else
#define FSB_TL(k) do { } while (0)
#endif

template <int MODE, int VARIANT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
flagstat_kernel_group(const uint16_t* __restrict__ base, uint64_t n,
                      unsigned long long* __restrict__ out, const __grid_constant__ XchgArgs xa)
{
    constexpr int DEPTH = 4;  // ring depth == group size
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (xa.pdl) pdl_launch_dependents();  // overlapped launches: the next kernel may take SM slots as they free up
    FSB_TL(0);

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
    // batches blockIdx.x, blockIdx.x + G, ... < NB = nb_q * G + nb_r (quotient and remainder from the host)
    const uint32_t my = (uint32_t)xa.nb_q + (blockIdx.x < xa.nb_r ? 1u : 0u);
    const uint64_t stride_bytes = G * (uint64_t)kVecPerBatch * 16u;

    GroupLanes<MODE, VARIANT> st;
    st.clear();
    uint32_t groups = 0;  // groups absorbed this epoch
    unsigned long long acc_all = 0ull, acc_fail = 0ull;

    // the CTA that would own batch NB takes the left-over vectors and the ragged
    // records as a partial group of two batches
    if (blockIdx.x == xa.nb_r) {
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
        st.template step<0>(w);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
        if (tid < head) w[0] = base[tid];
        else if (tid - head < ntail) w[0] = base[tail_start + (tid - head)];
        st.template step<1>(w);
        st.close();
        groups = 1;
    }

    // this thread's 16 bytes of load u of stage s:  smem + ((s * kU + u) * kThreads + tid) * 16
    const uint32_t my_smem = smem_u32(smem_raw) + tid * 16u;
    const unsigned char* src =
        reinterpret_cast<const unsigned char*>(body + tid + (uint64_t)blockIdx.x * kVecPerBatch);
    uint32_t fetched = 0;  // batches requested so far

    auto fetch = [&](auto stage_tag) {
        constexpr int S = decltype(stage_tag)::value;
        if (fetched < my) {
            cp_async16_imm<(S * kU + 0) * kThreads * 16, 0 * kThreads * 16>(my_smem, src);
            cp_async16_imm<(S * kU + 1) * kThreads * 16, 1 * kThreads * 16>(my_smem, src);
            cp_async16_imm<(S * kU + 2) * kThreads * 16, 2 * kThreads * 16>(my_smem, src);
            cp_async16_imm<(S * kU + 3) * kThreads * 16, 3 * kThreads * 16>(my_smem, src);
            src += stride_bytes;
            ++fetched;
        }
        cp_async_commit();  // empty groups keep the group count uniform
    };
    auto batch = [&](auto pos_tag, bool dense) -> bool {
        constexpr int POS = decltype(pos_tag)::value;
        cp_async_wait<DEPTH - 1>();
        uint32_t w[16];
        lds128_imm<(POS * kU + 0) * kThreads * 16>(my_smem, w[0], w[1], w[2], w[3]);
        lds128_imm<(POS * kU + 1) * kThreads * 16>(my_smem, w[4], w[5], w[6], w[7]);
        lds128_imm<(POS * kU + 2) * kThreads * 16>(my_smem, w[8], w[9], w[10], w[11]);
        lds128_imm<(POS * kU + 3) * kThreads * 16>(my_smem, w[12], w[13], w[14], w[15]);
        const bool general = st.template step<POS>(w, dense);
        fetch(pos_tag);  // the registers have been consumed: refill this stage
        return general;
    };
    using P0 = std::integral_constant<int, 0>;
    using P1 = std::integral_constant<int, 1>;
    using P2 = std::integral_constant<int, 2>;
    using P3 = std::integral_constant<int, 3>;

    fetch(P0{});
    fetch(P1{});
    fetch(P2{});
    fetch(P3{});
    FSB_TL(1);
#ifdef FSB_TIMELINE
    cp_async_wait<DEPTH - 1>();  // (probe only) when does the first stage land?
    FSB_TL(2);
#endif

    // One flush site: every pass of the outer loop is one epoch (at most kMaxGroups
    // groups); the ragged last group (my % 4 batches) rides in the last epoch with room.
    const uint32_t ngroups = my >> 2, rem = my & 3u;
    uint32_t g = 0;
    uint32_t dense_left = 0u;  // groups left to run without the OR-detect (warp-uniform)
    bool rem_done = rem == 0u;
    do {
        uint32_t lim = GCounter::kMaxGroups - groups;
        if (lim > ngroups - g) lim = ngroups - g;
        for (uint32_t i = 0; i < lim; ++i) {
            // Dense mode (VARIANT 3): a group whose four batches all needed the general
            // path switches the detect off for the next kDenseGroups groups.
            const bool dense = VARIANT == 3 && dense_left != 0u;
            const bool g0 = batch(P0{}, dense);
            const bool g1 = batch(P1{}, dense);
            const bool g2 = batch(P2{}, dense);
            const bool g3 = batch(P3{}, dense);
            if (VARIANT == 3) {
                if (dense) --dense_left;
                else if (g0 && g1 && g2 && g3) dense_left = kDenseGroups;
            }
        }
        g += lim;
        groups += lim;
        if (g == ngroups && !rem_done && groups < GCounter::kMaxGroups) {
            batch(P0{}, false);
            if (rem > 1) batch(P1{}, false);
            if (rem > 2) batch(P2{}, false);
            st.close();
            rem_done = true;
        }
        FSB_TL(3);
        acc_all += st.all.flush_warp(lane);
        if (MODE != kPospopcnt && st.fail_dirty) acc_fail += st.fail.flush_warp(lane);
        st.clear();
        groups = 0;
    } while (g < ngroups || !rem_done);
    cp_async_wait<0>();
    FSB_TL(4);

    cta_epilogue<MODE>(out, acc_all, acc_fail, n, warp, lane, true, xa);
    FSB_TL(5);
}

}  // namespace fsb200
