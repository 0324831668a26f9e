// lz4_block_group.cuh -- LZ4 block decoder, 32 sequences per warp step.
//
// lz4_block.cuh decodes one sequence per warp step: every lane parses the same
// token and the two copies are split over the lanes.  A block of FLAG words is
// ~10^5 sequences of a few bytes each (token, 1-2 literal records, a short
// match), so that kernel is bound by the ~100 dependent instructions of one
// sequence, not by bytes: ~700 cycles per sequence, 31 of 32 lanes idle.
//
// Here the warp works on a GROUP of up to 32 "simple" sequences at a time
// (literal length < 15, match length < 274: at most ONE match-length extension
// byte, so a sequence is 3 or 4 + lit bytes of input):
//
//   1. stage   256 bytes of input into shared memory; every lane also writes, for
//              its 8 byte positions p, nx[0][p] = p + size of a simple sequence
//              whose token would sit at p (p itself = not simple / does not fit /
//              last sequence of the block)
//   2. chase   by doubling: nx[L][p] = nx[L-1][nx[L-1][p]] for all 256 positions
//              (4 levels), then lane k walks k steps from the cursor with at most 5
//              dependent loads: the starts of the next K <= 32 sequences
//   3. parse   lane k decodes sequence k; a warp scan of lit + match lengths
//              gives every sequence its output position
//   4. copy    literals lane-parallel.  Matches of FLAG data mostly copy bytes that
//              another match of the SAME group produces (offsets of a few hundred
//              bytes), so every match byte first gets a parent pointer P[b] = the
//              output byte it copies (bytes produced before the group are copied at
//              once and become roots); pointer jumping P[b] = P[P[b]] over all bytes
//              of the group in parallel resolves the chains in log(depth) steps,
//              and a last pass copies root -> byte
//   5. flush   output goes to the shared-memory window only; whole 16-byte
//              chunks are written to HBM with coalesced uint4 stores
//
// Anything else -- literal runs >= 15, matches >= 274 bytes, the
// last sequence of a block, malformed input -- takes ONE step of the
// warp-cooperative decoder of lz4_block.cuh and returns to the group path.
// Same block format, same error codes, same Lz4BlockDesc interface.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lz4_block.cuh"

namespace fsb200 {

constexpr uint32_t kGrpWin = 256;      // input bytes examined per group (8 per lane)
constexpr uint32_t kGrpRing = 8192;    // bytes of a block's most recent output mirrored in shared memory
constexpr uint32_t kGrpOutCap = 2048;  // output bytes of one group = entries of the parent table P[]
// ring | win | nx[5][kGrpWin] | P[kGrpOutCap] u16
constexpr uint32_t kGrpSmemPerWarp = kGrpRing + 6u * kGrpWin + 2u * kGrpOutCap;
constexpr size_t kLz4GroupSmem = (size_t)kLz4WarpsPerCta * kGrpSmemPerWarp;

// ring -> global for output bytes [flushed, upto); without `exact` only up to the last
// 16-byte boundary of the global address (the rest waits for the next flush)
__device__ __forceinline__ void lz4g_flush(const Lz4Out& o, uint32_t& flushed, uint32_t upto, bool exact,
                                           uint32_t lane)
{
    uint32_t hi = upto;
    if (!exact) {
        const uint32_t a = (o.ga + upto) & ~15u;
        hi = a > o.ga ? a - o.ga : 0u;
    }
    if (hi <= flushed) return;
    uint32_t lo = flushed;
    const uint32_t mis = (o.ga + lo) & 15u;
    if (mis) {  // head: up to the next 16-byte boundary
        uint32_t hb = 16u - mis;
        if (hb > hi - lo) hb = hi - lo;
        if (lane < hb) o.g[lo + lane] = o.ring[o.ridx(lo + lane)];
        lo += hb;
    }
    const uint32_t nv = (hi - lo) >> 4;
    for (uint32_t v = lane; v < nv; v += 32u) {
        const uint32_t p = lo + (v << 4);
        *reinterpret_cast<uint4*>(o.g + p) = *reinterpret_cast<const uint4*>(o.ring + o.ridx(p));
    }
    lo += nv << 4;
    if (lane < hi - lo) o.g[lo + lane] = o.ring[o.ridx(lo + lane)];  // tail (exact only)
    flushed = hi;
}

// One sequence with the warp-cooperative copies of lz4_block.cuh (writes global AND ring).
// Requires every output byte below op to be in global memory.  Returns 0 = go on,
// 1 = block finished, < 0 = malformed.
__device__ __forceinline__ int lz4g_slow_sequence(const uint8_t* __restrict__ in, uint32_t in_size,
                                                  const Lz4Out& o, uint32_t out_cap, uint32_t lane,
                                                  uint32_t& ip, uint32_t& op)
{
    const uint32_t token = in[ip++];
    uint32_t lit = token >> 4;
    if (lit == 15u) {
        uint32_t b;
        do {
            if (ip >= in_size) return -1;
            b = in[ip++];
            lit += b;
        } while (b == 255u);
    }
    if (lit > in_size - ip || lit > out_cap - op) return -2;
    const uint32_t lit_at = ip;
    ip += lit;
    const bool last = ip >= in_size;  // the last sequence has no match part
    uint32_t offset = 0u, ml = 0u;
    if (!last) {
        if (in_size - ip < 2u) return -3;
        offset = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8);
        ip += 2;
        ml = token & 15u;
        if (ml == 15u) {
            uint32_t b;
            do {
                if (ip >= in_size) return -1;
                b = in[ip++];
                ml += b;
            } while (b == 255u);
        }
        ml += 4u;
        if (offset == 0u || offset > op + lit || ml > out_cap - op - lit) return -4;
    }
    warp_literals(o, op, in + lit_at, lit, lane);
    op += lit;
    if (last) return 1;
    __syncwarp();
    if (offset + ml <= o.mask + 1u) warp_match<true>(o, op, offset, ml, lane);
    else warp_match<false>(o, op, offset, ml, lane);
    op += ml;
    __syncwarp();
    return ip < in_size ? 0 : 1;
}

// Phase profile (tools/lz4_phase_probe.py builds with -DFSB_LZ4_PROFILE; never in the product):
// lane 0 of every warp sums clock64() deltas per phase of the group step.
#ifdef FSB_LZ4_PROFILE
__device__ unsigned long long g_lz4_prof[16];
#define LZ4P_DECL long long lz4p_t = clock64(); long long lz4p_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define LZ4P(k)                                   \
    do {                                          \
        const long long now_ = clock64();         \
        lz4p_acc[k] += now_ - lz4p_t;             \
        lz4p_t = now_;                            \
    } while (0)
#define LZ4P_COUNT(k, v) lz4p_acc[k] += (v)
#define LZ4P_DUMP                                                                       \
    do {                                                                                \
        if (lane == 0)                                                                  \
            for (int k_ = 0; k_ < 12; ++k_) atomicAdd(&g_lz4_prof[k_], (unsigned long long)lz4p_acc[k_]); \
    } while (0)
#else
#define LZ4P_DECL do { } while (0)
#define LZ4P(k) do { } while (0)
#define LZ4P_COUNT(k, v) do { } while (0)
#define LZ4P_DUMP do { } while (0)
#endif

// Returns the number of bytes produced, or a negative code for a malformed block.
__device__ __forceinline__ int lz4_decode_block_group(const uint8_t* __restrict__ in, uint32_t in_size,
                                                       const Lz4Out& o, uint8_t* win, uint8_t* nx,
                                                       uint16_t* P, uint32_t out_cap, uint32_t lane)
{
    constexpr uint32_t kFull = 0xffffffffu;
    uint32_t ip = 0u, op = 0u, flushed = 0u;
    if (in_size == 0u) return 0;
    LZ4P_DECL;
    for (;;) {
        LZ4P(9);  // loop overhead / flush tail of the previous step
        const uint32_t avail = in_size - ip;  // > 0
        uint32_t K = 0u, pos = 0u, endp = 0u;
        if (lane < 8u) {  // the next windows: have their lines on the way
            const uint32_t pf = ip + kGrpWin + lane * 128u;
            if (pf < in_size) asm volatile("prefetch.global.L1 [%0];" ::"l"(in + pf));
        }
        const uint32_t t0 = in[ip];
        bool simple = (t0 >> 4) != 15u;  // cheap look at the first token before staging a window
        if (simple && (t0 & 15u) == 15u) {
            const uint32_t e = ip + 3u + (t0 >> 4);
            simple = e < in_size && in[e] != 255u;
        }
        if (simple) {
            // 1. stage the window and the per-position sequence sizes
            uint8_t b[8];
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) {
                const uint32_t p = lane + 32u * j;
                b[j] = p < avail ? in[ip + p] : (uint8_t)0;
            }
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) win[lane + 32u * j] = b[j];
            __syncwarp();
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) {
                const uint32_t p = lane + 32u * j;
                const uint32_t lit = b[j] >> 4, mln = b[j] & 15u;
                uint32_t len = lit != 15u ? 3u + lit : 0u;
                if (mln == 15u && len != 0u) {  // one extension byte, and it must end the length
                    len += 1u;
                    if (p + len >= kGrpWin || win[p + len - 1u] == 255u) len = 0u;
                }
                // whole sequence inside the window (next position < 256: one byte), and another
                // token after it (a sequence that ends the input is the block's last one:
                // literals only, slow path)
                if (p + len >= kGrpWin || p + len >= avail) len = 0u;
                nx[p] = (uint8_t)(p + len);  // successor; a position that cannot start a sequence loops
            }
            __syncwarp();
            LZ4P(0);  // stage + successor sizes
            // 2. successor tables by doubling: nx[L][p] = 2^L-th successor of p
#pragma unroll
            for (uint32_t L = 1; L < 5u; ++L) {
                const uint8_t* a = nx + (L - 1u) * kGrpWin;
#pragma unroll
                for (uint32_t j = 0; j < 8u; ++j) {
                    const uint32_t p = lane + 32u * j;
                    nx[L * kGrpWin + p] = a[a[p]];
                }
                __syncwarp();
            }
            // lane k walks k steps from the cursor: binary decomposition, 5 dependent loads
            uint32_t p = 0u;
#pragma unroll
            for (uint32_t L = 0; L < 5u; ++L)
                if (lane & (1u << L)) p = nx[L * kGrpWin + p];
            pos = p;
            endp = nx[p];
            // a lane that landed on a looping position, or walked through one, is past the end
            K = __popc(__ballot_sync(kFull, endp != p));
            LZ4P(1);  // doubling + walk
        }
        if (K == 0u) {
            lz4g_flush(o, flushed, op, true, lane);
            const int r = lz4g_slow_sequence(in, in_size, o, out_cap, lane, ip, op);
            flushed = op;
            LZ4P(8);  // slow-path sequence
            LZ4P_COUNT(10, 1);  // slow sequences
            if (r < 0) return r;
            if (r == 1) break;
            continue;
        }

        // 3. parse: lane k owns sequence k
        bool mine = lane < K;
        const uint32_t pk = mine ? pos : 0u;
        const uint32_t tok = win[pk];
        const uint32_t litn = tok >> 4;
        const bool ext = (tok & 15u) == 15u;  // one match-length extension byte
        uint32_t lit = mine ? litn : 0u;
        uint32_t ml = mine ? (tok & 15u) + 4u + (ext ? (uint32_t)win[pk + 3u + litn] : 0u) : 0u;
        const uint32_t off = (uint32_t)win[pk + 1u + litn] | ((uint32_t)win[pk + 2u + litn] << 8);
        uint32_t tot = lit + ml;
        uint32_t incl = tot;
#pragma unroll
        for (uint32_t s = 1u; s < 32u; s <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, s);
            if (lane >= s) incl += t;
        }
        // P[] holds kGrpOutCap bytes: cut the group there (one sequence is at most 287 bytes)
        const uint32_t fit = __popc(__ballot_sync(kFull, mine && incl <= kGrpOutCap));
        if (fit < K) {
            K = fit;
            mine = lane < K;
            if (!mine) lit = ml = tot = 0u;
        }
        const uint32_t total = __shfl_sync(kFull, incl, (int)(K - 1u));
        const uint32_t consumed = __shfl_sync(kFull, endp, (int)(K - 1u));
        if (total > out_cap - op) return -4;
        const uint32_t o_k = op + incl - tot;  // first output byte of the sequence (mine only)
        const uint32_t m_k = o_k + lit;        // first byte of its match
        if (__any_sync(kFull, mine && (off == 0u || off > m_k))) return -4;
        LZ4P(2);  // parse + scan
        LZ4P_COUNT(7, 1);  // group steps; slot 6 sums their sequences
        LZ4P_COUNT(6, K);

        // 4a. literals (input never aliases output); a literal byte is its own root
        const uint32_t maxlit = __reduce_max_sync(kFull, lit);
        for (uint32_t i = 0; i < maxlit; ++i)
            if (i < lit) {
                o.ring[o.ridx(o_k + i)] = win[pk + 1u + i];
                P[o_k - op + i] = (uint16_t)(o_k - op + i);
            }

        LZ4P(3);  // literals
        // 4b. matches.  Every match byte gets a parent: the output byte it copies
        // (src + i, or src + i % off for an overlapping match, so that a run costs one
        // hop).  A parent produced before this group is copied right away and the byte
        // becomes a root; the others are resolved by pointer jumping over P[].
        const uint32_t src = m_k - off;
        const uint32_t gend = op + total;
        {
            // short matches lane-parallel (the lane walks its own bytes) ...
            const uint32_t n = __reduce_max_sync(kFull, (mine && !ext) ? ml : 0u);
            uint32_t r = 0u;
            for (uint32_t i = 0; i < n; ++i) {
                if (mine && !ext && i < ml) {
                    const uint32_t sp = src + r;  // absolute position of the parent
                    const uint32_t rel = m_k - op + i;
                    if (sp < op) {
                        o.ring[o.ridx(m_k + i)] = (gend - sp <= kGrpRing) ? o.ring[o.ridx(sp)] : o.g[sp];
                        P[rel] = (uint16_t)rel;
                    } else {
                        P[rel] = (uint16_t)(sp - op);
                    }
                    if (++r == off) r = 0u;
                }
            }
            // ... extended ones (19..273 bytes) 32 bytes per step by the whole warp
            uint32_t lm = __ballot_sync(kFull, mine && ext);
            while (lm) {
                const int j = __ffs((int)lm) - 1;
                lm &= lm - 1u;
                const uint32_t jm = __shfl_sync(kFull, m_k, j), js = __shfl_sync(kFull, src, j);
                const uint32_t jo = __shfl_sync(kFull, off, j), jn = __shfl_sync(kFull, ml, j);
                const uint32_t step = 32u % jo;
                uint32_t rr = lane % jo;
                for (uint32_t i = lane; i < jn; i += 32u) {
                    const uint32_t sp = js + rr;
                    const uint32_t rel = jm - op + i;
                    if (sp < op) {
                        o.ring[o.ridx(jm + i)] = (gend - sp <= kGrpRing) ? o.ring[o.ridx(sp)] : o.g[sp];
                        P[rel] = (uint16_t)rel;
                    } else {
                        P[rel] = (uint16_t)(sp - op);
                    }
                    rr += step;
                    if (rr >= jo) rr -= jo;
                }
            }
        }
        __syncwarp();
        LZ4P(4);  // parent pointers (short + extended matches)
        // pointer jumping, in place: whatever a lane reads from P[] is an ancestor
        for (;;) {
            bool changed = false;
            for (uint32_t b = lane; b < total; b += 32u) {
                const uint32_t p = P[b];
                if (p != b) {
                    const uint32_t q = P[p];
                    if (q != p) {
                        P[b] = (uint16_t)q;
                        changed = true;
                    }
                }
            }
            __syncwarp();
            if (!__any_sync(kFull, changed)) break;
        }
        // every match byte that is not a root copies its root (roots are never written here)
        for (uint32_t b = lane; b < total; b += 32u) {
            const uint32_t r = P[b];
            if (r != b) o.ring[o.ridx(op + b)] = o.ring[o.ridx(op + r)];
        }
        __syncwarp();
        LZ4P(5);  // pointer jumping + root copy
        ip += consumed;
        op += total;
        // 5. whole 16-byte chunks to HBM
        lz4g_flush(o, flushed, op, false, lane);
    }
    lz4g_flush(o, flushed, op, true, lane);
    LZ4P(9);
    LZ4P_DUMP;
    return (int)op;
}

// status[b] = decoded size (must equal raw_size) or a negative error code
__global__ void __launch_bounds__(kLz4WarpsPerCta * 32)
lz4_decode_group_kernel(const uint8_t* __restrict__ comp, uint8_t* raw, const Lz4BlockDesc* __restrict__ desc,
                        int* __restrict__ status, uint32_t n_blocks)
{
    extern __shared__ __align__(16) unsigned char lz4_smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wic = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * kLz4WarpsPerCta + wic;
    const uint32_t n_warps = gridDim.x * kLz4WarpsPerCta;
    unsigned char* mine = lz4_smem + wic * kGrpSmemPerWarp;
    for (uint32_t b = warp; b < n_blocks; b += n_warps) {
        const Lz4BlockDesc d = desc[b];
        Lz4Out o;
        o.g = raw + d.raw_off;
        o.ring = mine;
        o.ga = (uint32_t)(reinterpret_cast<uintptr_t>(o.g) & 15u);
        o.mask = kGrpRing - 1u;
        const int r = lz4_decode_block_group(
            comp + d.comp_off, d.comp_size, o, mine + kGrpRing, mine + kGrpRing + kGrpWin,
            reinterpret_cast<uint16_t*>(mine + kGrpRing + 6u * kGrpWin), d.raw_size, lane);
        if (lane == 0) status[b] = r;
        __syncwarp();
    }
}

}  // namespace fsb200
