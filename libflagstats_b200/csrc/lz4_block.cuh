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

// Returns the number of bytes produced, or a negative code for a malformed block.
__device__ __forceinline__ int lz4_decode_block_warp(const uint8_t* __restrict__ in, uint32_t in_size,
                                                      uint8_t* out, uint32_t out_cap, uint32_t lane)
{
    uint32_t ip = 0, op = 0;
    while (ip < in_size) {
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
        for (uint32_t i = lane; i < lit; i += 32u) out[op + i] = in[ip + i];
        ip += lit;
        op += lit;
        if (ip >= in_size) break;  // the last sequence has no match part
        if (in_size - ip < 2u) return -3;
        const uint32_t offset = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8);
        ip += 2;
        uint32_t ml = token & 15u;
        if (ml == 15u) {
            uint32_t b;
            do {
                if (ip >= in_size) return -1;
                b = in[ip++];
                ml += b;
            } while (b == 255u);
        }
        ml += 4u;
        if (offset == 0u || offset > op || ml > out_cap - op) return -4;
        __syncwarp();  // the literals (and everything before them) are visible to every lane
        const uint8_t* src = out + (op - offset);
        if (offset >= ml) {
            for (uint32_t i = lane; i < ml; i += 32u) out[op + i] = src[i];
        } else {  // overlapping match: periodic extension of the last `offset` bytes
            for (uint32_t i = lane; i < ml; i += 32u) out[op + i] = src[i % offset];
        }
        op += ml;
        __syncwarp();
    }
    return (int)op;
}

// status[b] = decoded size (must equal raw_size) or a negative error code
__global__ void __launch_bounds__(kLz4WarpsPerCta * 32)
lz4_decode_kernel(const uint8_t* __restrict__ comp, uint8_t* __restrict__ raw,
                  const Lz4BlockDesc* __restrict__ desc, int* __restrict__ status, uint32_t n_blocks)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = blockIdx.x * kLz4WarpsPerCta + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kLz4WarpsPerCta;
    for (uint32_t b = warp; b < n_blocks; b += n_warps) {
        const Lz4BlockDesc d = desc[b];
        const int r = lz4_decode_block_warp(comp + d.comp_off, d.comp_size, raw + d.raw_off, d.raw_size, lane);
        if (lane == 0) status[b] = r;
    }
}

}  // namespace fsb200
