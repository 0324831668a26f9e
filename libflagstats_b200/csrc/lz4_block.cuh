// lz4_block.cuh -- LZ4 *block format* decoder on the GPU, for the reference's
// compressed FLAG containers (benchmark/flagstats.cpp:110-186 writes
// [int32 raw_size][int32 comp_size][LZ4 block] records; :288-358 reads them with
// LZ4_decompress_safe and feeds every block to the flagstat kernel).
//
// lz4 itself is a system library the reference links (no version pinned, not
// vendored); what is restated here is the published block format: a block is a
// list of sequences
//     token (hi nibble = literal length, lo nibble = match length - 4)
//     [literal-length extension bytes: add 255 while the byte is 255]
//     literals
//     little-endian 16-bit match offset (1..65535, counted back from the output cursor)
//     [match-length extension bytes]
// and the last sequence ends after its literals.  A match may overlap the bytes
// it produces (offset < length), which repeats the last `offset` bytes.
//
// Mapping: ONE WARP PER BLOCK.  The sequence chain of a block is inherently
// serial, so all lanes parse the token stream redundantly (uniform loads) and
// split the two copies of every sequence across the 32 lanes; blocks are
// independent, and a file has thousands of them, which is where the
// parallelism comes from.  The decoded records are written to HBM exactly once
// and then read exactly once by the flagstat kernel; the compressed bytes are
// what crosses PCIe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fsb200 {

struct Lz4BlockDesc {
    unsigned long long comp_off;  // byte offset of the block's payload in the compressed buffer
    unsigned long long raw_off;   // byte offset of its output in the decoded buffer
    uint32_t comp_size;
    uint32_t raw_size;            // expected decoded size (the container's header field)
};

constexpr int kLz4WarpsPerCta = 4;
constexpr uint32_t kLz4Win = 16384;  // bytes of a block's most recent output mirrored in shared memory
constexpr uint32_t kLz4Unroll = 8;   // independent byte loads in flight per lane and copy step
constexpr size_t kLz4Smem = (size_t)kLz4WarpsPerCta * kLz4Win;

// One warp's view of the block it is decoding: global output + a shared-memory ring that
// mirrors the last kLz4Win bytes.  Short matches -- the latency-critical ones: a block of
// FLAG words is ~10^5 sequences of a few bytes each -- read their source from the ring
// (~30 cycles) instead of from L2 (~300 cycles: global stores are not kept in L1).
struct Lz4Out {
    uint8_t* g;      // global output of this block
    uint8_t* ring;   // this warp's kLz4Win bytes of shared memory (16-byte aligned)
    uint32_t ga;     // (address of g) & 15: the ring is indexed so that 16-byte chunks of the
                     // ring and of global memory line up (lz4_block_group.cuh flushes in uint4)
    uint32_t mask;   // ring bytes - 1 (kLz4Win here, kGrpRing in lz4_block_group.cuh)
    __device__ __forceinline__ uint32_t ridx(uint32_t pos) const { return (pos + ga) & mask; }
    __device__ __forceinline__ void put(uint32_t pos, uint8_t v) const
    {
        g[pos] = v;
        ring[ridx(pos)] = v;
    }
};

// out[op + i] = src[i], i < n: literals from the compressed stream (never aliases the output)
__device__ __forceinline__ void warp_literals(const Lz4Out& o, uint32_t op, const uint8_t* __restrict__ src,
                                              uint32_t n, uint32_t lane)
{
    uint32_t i = lane;
    for (; i + 32u * (kLz4Unroll - 1u) < n; i += 32u * kLz4Unroll) {
        uint8_t v[kLz4Unroll];
#pragma unroll
        for (uint32_t u = 0; u < kLz4Unroll; ++u) v[u] = src[i + 32u * u];
#pragma unroll
        for (uint32_t u = 0; u < kLz4Unroll; ++u) o.put(op + i + 32u * u, v[u]);
    }
    for (; i < n; i += 32u) o.put(op + i, src[i]);
}

// out[op + i] = out[op - offset + (i % offset)], i < n.  For offset >= n that is a plain copy,
// for offset < n the LZ4 overlapping match (periodic extension of the last `offset` bytes).
// All source bytes were produced by earlier sequences (the caller has synchronised the warp).
// FROM_RING requires offset + n <= kLz4Win: then no byte written here reuses a ring slot
// that still holds a source byte.
template <bool FROM_RING>
__device__ __forceinline__ void warp_match(const Lz4Out& o, uint32_t op, uint32_t offset, uint32_t n,
                                           uint32_t lane)
{
    const uint32_t base = op - offset;
    if (offset >= n) {  // plain copy: no index arithmetic modulo the period
        uint32_t i = lane;
        for (; i + 32u * (kLz4Unroll - 1u) < n; i += 32u * kLz4Unroll) {
            uint8_t v[kLz4Unroll];
#pragma unroll
            for (uint32_t u = 0; u < kLz4Unroll; ++u) {
                const uint32_t p = base + i + 32u * u;
                v[u] = FROM_RING ? o.ring[o.ridx(p)] : o.g[p];
            }
#pragma unroll
            for (uint32_t u = 0; u < kLz4Unroll; ++u) o.put(op + i + 32u * u, v[u]);
        }
        for (; i < n; i += 32u) {
            const uint32_t p = base + i;
            o.put(op + i, FROM_RING ? o.ring[o.ridx(p)] : o.g[p]);
        }
        return;
    }
    const uint32_t step = 32u % offset;  // a lane's phase advances by this (mod offset) per 32 bytes
    uint32_t r = lane % offset;
    uint32_t i = lane;
    for (; i + 32u * (kLz4Unroll - 1u) < n; i += 32u * kLz4Unroll) {
        uint8_t v[kLz4Unroll];
#pragma unroll
        for (uint32_t u = 0; u < kLz4Unroll; ++u) {
            v[u] = FROM_RING ? o.ring[o.ridx(base + r)] : o.g[base + r];
            r += step;
            if (r >= offset) r -= offset;
        }
#pragma unroll
        for (uint32_t u = 0; u < kLz4Unroll; ++u) o.put(op + i + 32u * u, v[u]);
    }
    for (; i < n; i += 32u) {
        const uint8_t v = FROM_RING ? o.ring[o.ridx(base + r)] : o.g[base + r];
        o.put(op + i, v);
        r += step;
        if (r >= offset) r -= offset;
    }
}

// Returns the number of bytes produced, or a negative code for a malformed block.
__device__ __forceinline__ int lz4_decode_block_warp(const uint8_t* __restrict__ in, uint32_t in_size,
                                                      const Lz4Out& o, uint32_t out_cap, uint32_t lane)
{
    uint32_t ip = 0, op = 0;
    if (in_size == 0u) return 0;
    uint32_t token = in[ip++];
    for (;;) {
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
        // the next token's address is known before any byte is copied: have it in flight
        const bool more = !last && ip < in_size;
        const uint32_t next_token = more ? in[ip] : 0u;

        if (!last && lit <= 32u && ml <= 32u) {
            // Fast path -- a block of FLAG words is ~10^5 sequences of a few bytes each, so the
            // serial chain is instruction-latency bound: one byte per lane, no loops, and the
            // match source comes from the shared-memory ring whenever it is near enough.
            if (lane < lit) o.put(op + lane, in[lit_at + lane]);
            op += lit;
            __syncwarp();  // the literals (and everything before them) are visible to every lane
            if (lane < ml) {
                uint32_t rel = lane;
                if (offset < ml) rel = lane % offset;  // warp-uniform branch: overlapping match
                const uint32_t sidx = op - offset + rel;
                const uint8_t v = (offset <= kLz4Win - 32u) ? o.ring[o.ridx(sidx)] : o.g[sidx];
                o.put(op + lane, v);
            }
            op += ml;
            __syncwarp();
        } else {
            warp_literals(o, op, in + lit_at, lit, lane);
            op += lit;
            if (last) break;
            __syncwarp();
            if (offset + ml <= kLz4Win) warp_match<true>(o, op, offset, ml, lane);
            else warp_match<false>(o, op, offset, ml, lane);
            op += ml;
            __syncwarp();
        }
        if (!more) break;
        token = next_token;
        ++ip;
    }
    return (int)op;
}

// status[b] = decoded size (must equal raw_size) or a negative error code
__global__ void __launch_bounds__(kLz4WarpsPerCta * 32)
lz4_decode_kernel(const uint8_t* __restrict__ comp, uint8_t* raw, const Lz4BlockDesc* __restrict__ desc,
                  int* __restrict__ status, uint32_t n_blocks)
{
    extern __shared__ __align__(16) unsigned char lz4_smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wic = threadIdx.x >> 5;
    const uint32_t warp = blockIdx.x * kLz4WarpsPerCta + wic;
    const uint32_t n_warps = gridDim.x * kLz4WarpsPerCta;
    for (uint32_t b = warp; b < n_blocks; b += n_warps) {
        const Lz4BlockDesc d = desc[b];
        Lz4Out o;
        o.g = raw + d.raw_off;
        o.ring = lz4_smem + wic * kLz4Win;
        o.ga = 0u;
        o.mask = kLz4Win - 1u;
        const int r = lz4_decode_block_warp(comp + d.comp_off, d.comp_size, o, d.raw_size, lane);
        if (lane == 0) status[b] = r;
        __syncwarp();
    }
}

}  // namespace fsb200
