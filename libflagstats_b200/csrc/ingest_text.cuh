// ingest_text.cuh -- FLAG ingest on the GPU: the text column `samtools view | cut -f 2`
// produces (one decimal FLAG per line) -> the uint16 FLAG column the hot path reads.
// Device twin of benchmark/utility.cpp:29-32:
//     while (std::getline(std::cin, str)) { uint16_t val = std::atoi(str.c_str()); write(val); }
// i.e. one record per '\n'-terminated line (a last line without '\n' counts too), atoi
// semantics (leading white space, optional sign, digits up to the first non-digit, 0 when
// there are none), truncated to 16 bits.
//
// Two passes over the text, 4 KiB tiles (256 threads x 16 bytes):
//   count   newlines per tile                       -> exclusive scan (CUB) -> first line index of each tile
//   parse   every thread owns the newlines inside its 16 bytes; for each it walks back to the
//           start of that line (lines are a few characters), parses forward, stores out[line]
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fsb200 {

constexpr int kIngestThreads = 256;
constexpr int kIngestBytesPerThread = 16;
constexpr int kIngestTile = kIngestThreads * kIngestBytesPerThread;

__device__ __forceinline__ uint32_t newline_mask16(const uint8_t* __restrict__ text, uint64_t n, uint64_t at)
{
    uint32_t m = 0;
    if (at + 16 <= n && ((reinterpret_cast<uintptr_t>(text) + at) & 15u) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + at);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (((w[k] >> (8 * b)) & 0xFFu) == (uint32_t)'\n') m |= 1u << (4 * k + b);
    } else {
        for (int i = 0; i < 16; ++i)
            if (at + i < n && text[at + i] == (uint8_t)'\n') m |= 1u << i;
    }
    return m;
}

__global__ void __launch_bounds__(kIngestThreads)
ingest_count_kernel(const uint8_t* __restrict__ text, uint64_t n, unsigned long long* __restrict__ tile_lines)
{
    const uint64_t at = (uint64_t)blockIdx.x * kIngestTile + (uint64_t)threadIdx.x * kIngestBytesPerThread;
    uint32_t c = at < n ? __popc(newline_mask16(text, n, at)) : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t s[kIngestThreads / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < kIngestThreads / 32; ++i) t += s[i];
        tile_lines[blockIdx.x] = t;
    }
}

// atoi() of the line [lo, hi): C locale white space, optional sign, decimal digits
__device__ __forceinline__ uint16_t atoi_u16(const uint8_t* __restrict__ text, uint64_t lo, uint64_t hi)
{
    uint64_t p = lo;
    while (p < hi) {
        const uint8_t ch = text[p];
        if (ch == ' ' || (ch >= 9 && ch <= 13)) ++p;
        else break;
    }
    bool neg = false;
    if (p < hi && (text[p] == '-' || text[p] == '+')) {
        neg = text[p] == '-';
        ++p;
    }
    uint32_t v = 0;
    while (p < hi) {
        const uint32_t d = (uint32_t)text[p] - (uint32_t)'0';
        if (d > 9u) break;
        v = v * 10u + d;
        ++p;
    }
    if (neg) v = 0u - v;
    return (uint16_t)v;
}

__global__ void __launch_bounds__(kIngestThreads)
ingest_parse_kernel(const uint8_t* __restrict__ text, uint64_t n,
                    const unsigned long long* __restrict__ tile_first_line, uint16_t* __restrict__ out)
{
    const uint64_t at = (uint64_t)blockIdx.x * kIngestTile + (uint64_t)threadIdx.x * kIngestBytesPerThread;
    uint32_t m = at < n ? newline_mask16(text, n, at) : 0u;
    const uint32_t mine = __popc(m);
    // exclusive scan of `mine` over the CTA
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    __shared__ uint32_t s[kIngestThreads / 32];
    if (lane == 31) s[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s[w];
    unsigned long long line = tile_first_line[blockIdx.x] + before + (incl - mine);
    while (m) {
        const uint32_t i = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t nl = at + i;  // this line ends here
        uint64_t lo = nl;
        while (lo > 0 && text[lo - 1] != (uint8_t)'\n') --lo;
        out[line++] = atoi_u16(text, lo, nl);
    }
}

}  // namespace fsb200
