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
//              its 8 byte positions p, len[p] = size of a simple sequence whose
//              token would sit at p (0 = not simple / does not fit / last sequence)
//   2. chase   lane 0 follows p += len[p] from the cursor: the starts of the next
//              K <= 32 sequences (the only serial step: one LDS per sequence)
//   3. parse   lane k decodes sequence k; a warp scan of lit + match lengths
//              gives every sequence its output position
//   4. copy    literals lane-parallel; matches in dependency rounds: a match is
//              copied once no still-pending earlier match of the group overlaps
//              its source (for FLAG runs the source is the sequence's own
//              literals: one round).  Within a round the short matches (< 19
//              bytes) are copied lane-parallel, the extended ones one after the
//              other by the whole warp.
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

constexpr uint32_t kGrpWin = 256;  // input bytes examined per group (8 per lane)
constexpr uint32_t kGrpSmemPerWarp = kLz4Win + 2u * kGrpWin + 64u;  // ring | win | len | pos[32] u16
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
    if (offset + ml <= kLz4Win) warp_match<true>(o, op, offset, ml, lane);
    else warp_match<false>(o, op, offset, ml, lane);
    op += ml;
    __syncwarp();
    return ip < in_size ? 0 : 1;
}

// ring[dst + i] = source[src + (i % offset)], i < n, by the whole warp; the source is the ring
// or (further back than the window) global memory.  Only the first min(n, offset) source
// bytes are read, all of them produced before this call.
__device__ __forceinline__ void lz4g_warp_match(const Lz4Out& o, uint32_t dst, uint32_t src, uint32_t offset,
                                                uint32_t n, bool from_ring, uint32_t lane)
{
    if (offset >= n) {
        for (uint32_t i = lane; i < n; i += 32u)
            o.ring[o.ridx(dst + i)] = from_ring ? o.ring[o.ridx(src + i)] : o.g[src + i];
        return;
    }
    const uint32_t step = 32u % offset;
    uint32_t r = lane % offset;
    for (uint32_t i = lane; i < n; i += 32u) {
        o.ring[o.ridx(dst + i)] = from_ring ? o.ring[o.ridx(src + r)] : o.g[src + r];
        r += step;
        if (r >= offset) r -= offset;
    }
}

// number of lanes j whose (ascending) value v_j is <= x  [STRICT: < x]; the answer must be < 32
template <bool STRICT>
__device__ __forceinline__ uint32_t lz4g_count_below(uint32_t v, uint32_t x)
{
    uint32_t cnt = 0u;
#pragma unroll
    for (uint32_t s = 16u; s >= 1u; s >>= 1) {
        const uint32_t t = __shfl_sync(0xffffffffu, v, (int)(cnt + s - 1u));
        if (STRICT ? (t < x) : (t <= x)) cnt += s;
    }
    return cnt;
}

// Returns the number of bytes produced, or a negative code for a malformed block.
__device__ __forceinline__ int lz4_decode_block_group(const uint8_t* __restrict__ in, uint32_t in_size,
                                                       const Lz4Out& o, uint8_t* win, uint8_t* lent,
                                                       uint16_t* posv, uint32_t out_cap, uint32_t lane)
{
    constexpr uint32_t kFull = 0xffffffffu;
    uint32_t ip = 0u, op = 0u, flushed = 0u;
    if (in_size == 0u) return 0;
    for (;;) {
        const uint32_t avail = in_size - ip;  // > 0
        uint32_t K = 0u, consumed = 0u;
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
                    if (p + len > kGrpWin || win[p + len - 1u] == 255u) len = 0u;
                }
                // whole sequence inside the window, and another token after it (a sequence
                // that ends the input is the block's last one: literals only, slow path)
                if (p + len > kGrpWin || p + len >= avail) len = 0u;
                lent[p] = (uint8_t)len;
            }
            __syncwarp();
            // 2. the serial step: starts of the next K sequences
            uint32_t p = 0u, k = 0u;
            if (lane == 0u) {
                while (k < 32u && p < kGrpWin) {
                    const uint32_t l = lent[p];
                    if (l == 0u) break;
                    posv[k++] = (uint16_t)p;
                    p += l;
                }
            }
            K = __shfl_sync(kFull, k, 0);
            consumed = __shfl_sync(kFull, p, 0);
            __syncwarp();
        }
        if (K == 0u) {
            lz4g_flush(o, flushed, op, true, lane);
            const int r = lz4g_slow_sequence(in, in_size, o, out_cap, lane, ip, op);
            flushed = op;
            if (r < 0) return r;
            if (r == 1) break;
            continue;
        }

        // 3. parse: lane k owns sequence k
        const bool mine = lane < K;
        const uint32_t pk = mine ? (uint32_t)posv[lane] : 0u;
        const uint32_t tok = win[pk];
        const uint32_t lit = mine ? (tok >> 4) : 0u;
        const bool ext = mine && (tok & 15u) == 15u;  // one match-length extension byte
        const uint32_t ml = mine ? (tok & 15u) + 4u + (ext ? (uint32_t)win[pk + 3u + (tok >> 4)] : 0u) : 0u;
        const uint32_t off = (uint32_t)win[pk + 1u + (tok >> 4)] | ((uint32_t)win[pk + 2u + (tok >> 4)] << 8);
        const uint32_t tot = lit + ml;
        uint32_t incl = tot;
#pragma unroll
        for (uint32_t s = 1u; s < 32u; s <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, s);
            if (lane >= s) incl += t;
        }
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        if (total > out_cap - op) return -4;
        const uint32_t o_k = op + incl - tot;  // first output byte of the sequence
        const uint32_t m_k = o_k + lit;        // first byte of its match
        if (__any_sync(kFull, mine && (off == 0u || off > m_k))) return -4;

        // 4a. literals (input never aliases output)
        const uint32_t maxlit = __reduce_max_sync(kFull, lit);
        for (uint32_t i = 0; i < maxlit; ++i)
            if (i < lit) o.ring[o.ridx(o_k + i)] = win[pk + 1u + i];

        // 4b. matches.  Sequence k reads [src, src + min(ml, off)) from other sequences (the
        // rest of an overlapping match is its own output).  It depends on exactly the
        // earlier sequences j whose match region [m_j, end_j) meets that range.
        const uint32_t src = m_k - off;
        const uint32_t src_hi = src + (ml < off ? ml : off);
        const uint32_t endv = mine ? op + incl : 0xffffffffu;
        const uint32_t mv = mine ? m_k : 0xffffffffu;
        const uint32_t jlo = lz4g_count_below<false>(endv, src);    // end_j <= src: entirely before
        const uint32_t jhi = lz4g_count_below<true>(mv, src_hi);    // m_j < src_hi
        const uint32_t dep = (mine && jhi > jlo) ? (((1u << jhi) - 1u) & ~((1u << jlo) - 1u)) : 0u;
        // The source is still in the ring unless it lies more than a window behind the end
        // of this group; then it is in global memory (flushed long ago) and cannot overlap.
        const bool from_ring = (op + total) - src <= kLz4Win;
        __syncwarp();  // literals visible
        bool pending = mine;
        for (;;) {
            const uint32_t pm = __ballot_sync(kFull, pending);
            if (pm == 0u) break;
            const bool go = pending && (pm & dep) == 0u;
            // short matches: every lane copies its own, byte by byte (an overlapping match
            // reads what the same lane wrote a few iterations earlier)
            const uint32_t n = __reduce_max_sync(kFull, (go && !ext) ? ml : 0u);
            if (from_ring) {
                for (uint32_t i = 0; i < n; ++i)
                    if (go && !ext && i < ml) o.ring[o.ridx(m_k + i)] = o.ring[o.ridx(src + i)];
            } else {
                for (uint32_t i = 0; i < n; ++i)
                    if (go && !ext && i < ml) o.ring[o.ridx(m_k + i)] = o.g[src + i];
            }
            // extended matches (19..273 bytes): one after the other, 32 bytes per step
            uint32_t lm = __ballot_sync(kFull, go && ext);
            while (lm) {
                const int j = __ffs((int)lm) - 1;
                lm &= lm - 1u;
                lz4g_warp_match(o, __shfl_sync(kFull, m_k, j), __shfl_sync(kFull, src, j),
                                __shfl_sync(kFull, off, j), __shfl_sync(kFull, ml, j),
                                __shfl_sync(kFull, (int)from_ring, j) != 0, lane);
            }
            if (go) pending = false;
            __syncwarp();
        }
        ip += consumed;
        op += total;
        // 5. whole 16-byte chunks to HBM
        lz4g_flush(o, flushed, op, false, lane);
    }
    lz4g_flush(o, flushed, op, true, lane);
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
        const int r = lz4_decode_block_group(comp + d.comp_off, d.comp_size, o, mine + kLz4Win,
                                             mine + kLz4Win + kGrpWin,
                                             reinterpret_cast<uint16_t*>(mine + kLz4Win + 2u * kGrpWin),
                                             d.raw_size, lane);
        if (lane == 0) status[b] = r;
        __syncwarp();
    }
}

}  // namespace fsb200
