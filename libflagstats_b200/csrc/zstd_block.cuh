// zstd_block.cuh -- device launch of the Zstd frame decoder (zstd_frame.cuh) over a batch of the
// reference's container records; same descriptor / status interface as the LZ4 decoders
// (lz4_block.cuh), so the block-file pipeline treats the two codecs alike.
//
// First version: one WARP slot per frame with lane 0 decoding (the decoder is serial code; one
// CTA = one warp = one frame lets the hardware spread the ~1600 frames of a file over all SMs,
// up to 32 frames per SM).  Tables, literal buffer and repeat offsets live in a per-frame
// workspace in global memory (zstd::Work, ~137 KiB).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lz4_block.cuh"
#include "lz4_block_cta.cuh"
#include "zstd_frame.cuh"

namespace fsb200 {

// status[b] = decoded size (must equal raw_size) or a negative zstd::kErr* code
__global__ void __launch_bounds__(32)
zstd_decode_kernel(const uint8_t* __restrict__ comp, uint8_t* raw, const Lz4BlockDesc* __restrict__ desc,
                   int* __restrict__ status, uint32_t n_blocks, zstd::Work* work)
{
    const uint32_t b = blockIdx.x;
    if (b >= n_blocks || threadIdx.x != 0) return;
    const Lz4BlockDesc d = desc[b];
    const int64_t r = zstd::decode_frame(comp + d.comp_off, d.comp_size, raw + d.raw_off, d.raw_size, work[b]);
    status[b] = (int)r;
}

// ---------------------------------------------------------------------------------------------
// Second version (the default): two kernels per batch of frames.
//   zstd_parse_kernel  entropy stage, ONE LANE per frame (the FSE / Huffman bit streams are serial), four
//                      frames per CTA, sixteen per SM: the tables of a frame (13.7 KiB) live in shared
//                      memory and the bit reader keeps 64 bits in registers, so a sequence costs a couple of
//                      shared-memory round trips instead of several global ones.  Output: 16-byte
//                      descriptors {output position, literal position, literal length, offset} and the
//                      frame's literal bytes, in global memory.
//   zstd_copy_kernel   one CTA (512 threads) per frame runs the copy phase of the LZ4 decoder (l4_copy,
//                      lz4_block_cta.cuh) over those descriptors: parents, pointer jumping, 8 KiB tiles.
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(zstd::SeqDesc) == sizeof(L4Desc), "parse_frame writes what l4_copy reads");

struct ZstdFrameAux {   // where a frame's slices of the batch buffers are (host-computed)
    uint64_t desc_off;  // in descriptors
    uint64_t lit_off;   // in bytes, 16-byte aligned
    uint32_t desc_cap;
    uint32_t lit_cap;
};
struct ZstdParsed {
    int status;         // bytes the descriptors produce, or a negative zstd::kErr* code
    uint32_t nd;
    uint32_t lit_used;
    uint32_t pad;
};

constexpr int kZstdFramesPerCta = 4;
constexpr size_t kZstdParseSmem = sizeof(zstd::Tables) * kZstdFramesPerCta;

__global__ void __launch_bounds__(32 * kZstdFramesPerCta)
zstd_parse_kernel(const uint8_t* __restrict__ comp, const Lz4BlockDesc* __restrict__ desc,
                  const ZstdFrameAux* __restrict__ aux, uint32_t n_blocks, uint8_t* lit, L4Desc* descs,
                  ZstdParsed* parsed)
{
    extern __shared__ __align__(16) unsigned char zstd_smem[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t b = blockIdx.x * kZstdFramesPerCta + warp;
    if (b >= n_blocks || (threadIdx.x & 31u) != 0u) return;
    zstd::Tables& t = reinterpret_cast<zstd::Tables*>(zstd_smem)[warp];
    const Lz4BlockDesc d = desc[b];
    const ZstdFrameAux a = aux[b];
    uint32_t nd = 0;
    uint64_t lit_used = 0;
    const int64_t r = zstd::parse_frame(comp + d.comp_off, d.comp_size, d.raw_size, t, lit + a.lit_off, a.lit_cap,
                                        reinterpret_cast<zstd::SeqDesc*>(descs + a.desc_off), a.desc_cap, &nd, &lit_used);
    parsed[b] = ZstdParsed{(int)r, nd, (uint32_t)lit_used, 0u};
}

// status[b] = decoded size (must equal raw_size) or a negative code.  tf: gridDim.x regions of tf_stride words
// (the tile index of l4_copy).
__global__ void __launch_bounds__(kL4Threads, 2)
zstd_copy_kernel(uint8_t* raw, const Lz4BlockDesc* __restrict__ desc, const ZstdFrameAux* __restrict__ aux,
                 const ZstdParsed* parsed, int* __restrict__ status, uint32_t n_blocks, const uint8_t* lit, L4Desc* descs,
                 uint32_t* tf, uint32_t tf_stride)
{
    extern __shared__ __align__(16) unsigned char l4_smem[];
    uint8_t* ring = l4_smem;
    uint16_t* P = reinterpret_cast<uint16_t*>(l4_smem + kL4Ring);
    L4Shared* sh = reinterpret_cast<L4Shared*>(l4_smem + kL4Ring + 2u * kL4Tile);
    L4Long* longs = reinterpret_cast<L4Long*>(l4_smem + kL4Ring + 2u * kL4Tile + 256u);
    L4Desc* dsm = reinterpret_cast<L4Desc*>(l4_smem + kL4Ring + 2u * kL4Tile + 256u + sizeof(L4Long) * kL4MaxLong);
    uint32_t* tile_first = tf + (size_t)blockIdx.x * tf_stride;
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const Lz4BlockDesc d = desc[b];
        const ZstdFrameAux a = aux[b];
        const ZstdParsed ps = parsed[b];
        if (threadIdx.x == 0) sh->err = 0;
        __syncthreads();
        int r = ps.status;
        if (r > 0 && (uint32_t)r <= d.raw_size)
            r = l4_copy(lit + a.lit_off, ps.lit_used, raw + d.raw_off, (uint32_t)r, descs + a.desc_off, ps.nd, tile_first,
                        ring, P, sh, longs, dsm, 0xFFFFFFFFu);
        __syncthreads();
        if (threadIdx.x == 0) status[b] = r;
        __syncthreads();
    }
}

}  // namespace fsb200
