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

}  // namespace fsb200
