// bitcounter.cuh -- bit-sliced (vertical) counters for positional population
// counts on sm_100a.
//
// A BitCounter counts, for each of the 32 bit positions of a 32-bit word, how
// many of the words fed to it had that bit set.  The count of position j is
// stored "vertically": bit j of plane k holds bit k of the count.  Feeding a
// word is a carry-save addition, i.e. two LOP3 per full adder with the truth
// tables 0x96 (a^b^c) and 0xE8 (majority) -- the same two functions the
// reference's AVX-512 carry-save step uses (libalgebra/libalgebra.h:2311-2319)
// -- so the steady-state cost is ~2 LOP3 per input word however many of the 32
// positions are in use.
//
// Structure (chosen for the B200 integer pipe budget):
//   level 1   Harley-Seal over a batch of 16 words into planes 1,2,4,8 and one
//             carry word of weight 16                       (15 full adders)
//   hold      NHOLD binary-counter levels (weights 16,32,..) each with a plane
//             and a pending "hold" word: one full adder per level per 2^L
//             batches, selected by the (warp-uniform) batch index
//   up        NUP ripple planes above that (half adders, touched once every
//             2^NHOLD batches)
// Nothing is ever expanded to per-position integers inside the hot loop; that
// happens once per epoch in flush(), cooperatively across the warp.
#pragma once
#include <cstdint>

namespace fsb200 {

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// full adder: (hi, lo) = a + b + c, bitwise in all 32 positions
#define FSB_CSA(hi, lo, a, b, c)              \
    do {                                      \
        const uint32_t a_ = (a), b_ = (b), c_ = (c); \
        (hi) = maj3(a_, b_, c_);              \
        (lo) = xor3(a_, b_, c_);              \
    } while (0)

template <int NHOLD, int NUP>
struct BitCounter {
    static constexpr int kPlanes = 4 + NHOLD + NUP;          // normalised planes
    static constexpr uint32_t kMaxBatches = (1u << (NHOLD + NUP)) - 2u;  // even, < 2^(NHOLD+NUP)
    static_assert(NHOLD >= 1 && NUP >= 1, "need at least one hold and one up level");

    uint32_t p1, p2, p4, p8;
    uint32_t pl[NHOLD];
    uint32_t hd[NHOLD];
    uint32_t up[NUP];

    __device__ __forceinline__ void clear()
    {
        p1 = p2 = p4 = p8 = 0u;
#pragma unroll
        for (int i = 0; i < NHOLD; ++i) { pl[i] = 0u; hd[i] = 0u; }
#pragma unroll
        for (int i = 0; i < NUP; ++i) up[i] = 0u;
    }

    // Absorb 16 words.  b = number of batches absorbed since clear(); it must be
    // the same for every thread of the warp (it is a loop counter) and stay
    // below kMaxBatches + 2.
    __device__ __forceinline__ void absorb16(const uint32_t (&x)[16], uint32_t b)
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
        FSB_CSA(c, p8, p8, t8a, t8b);  // c has weight 16

        bool go = true;
#pragma unroll
        for (int L = 0; L < NHOLD; ++L) {
            if (go) {
                if (((b >> L) & 1u) == 0u) {
                    hd[L] = c;
                    go = false;
                } else {
                    FSB_CSA(c, pl[L], pl[L], hd[L], c);
                }
            }
        }
        if (go) {
#pragma unroll
            for (int k = 0; k < NUP; ++k) {
                const uint32_t t = up[k] & c;
                up[k] ^= c;
                c = t;
            }
        }
    }

    // Expand to plain planes of weight 2^k.  nb = batches absorbed since clear().
    __device__ __forceinline__ void normalize(uint32_t nb, uint32_t (&N)[kPlanes]) const
    {
        N[0] = p1; N[1] = p2; N[2] = p4; N[3] = p8;
#pragma unroll
        for (int L = 0; L < NHOLD; ++L) N[4 + L] = pl[L];
#pragma unroll
        for (int k = 0; k < NUP; ++k) N[4 + NHOLD + k] = up[k];
#pragma unroll
        for (int L = 0; L < NHOLD; ++L) {
            // hold L is pending iff bit L of nb is set (binary counter)
            uint32_t c = ((nb >> L) & 1u) ? hd[L] : 0u;
#pragma unroll
            for (int k = 4 + L; k < kPlanes; ++k) {
                const uint32_t t = N[k] & c;
                N[k] ^= c;
                c = t;
            }
        }
    }

    // Warp-cooperative expansion: returns, in lane j, the warp-wide total of bit
    // position j.  All 32 lanes must call it.  Bit-sliced butterfly add over the
    // 5 lane-exchange steps (planes grow by one per step), then each lane reads
    // its own column.
    __device__ __forceinline__ uint32_t flush_warp(uint32_t nb, uint32_t lane) const
    {
        uint32_t N[kPlanes + 5];
        {
            uint32_t M[kPlanes];
            normalize(nb, M);
#pragma unroll
            for (int k = 0; k < kPlanes; ++k) N[k] = M[k];
        }
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

}  // namespace fsb200
