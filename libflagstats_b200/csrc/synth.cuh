// synth.cuh -- device twins of the oracle's synthetic FLAG generators
// (oracle/flagstat_oracle.c: oracle_synth_uniform / oracle_synth_hiseqx).
// Pure functions of the GLOBAL record index, so any shard on any GPU and the
// CPU oracle regenerate identical columns without moving data
// (SURVEY.md section 8d).  Test / bench support, not part of the hot path.
#pragma once
#include <cstdint>

namespace fsb200 {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// U(0, mask): four records per hash -- the distribution of
// benchmark/generate.cpp:11 and benchmark/inmemory.cpp:113 when mask = 0x0FFF.
__host__ __device__ __forceinline__ uint16_t synth_uniform_at(uint64_t seed, uint64_t i,
                                                             uint16_t mask)
{
    const uint64_t h = mix64(seed + ((i >> 2) + 1ull) * 0x9E3779B97F4A7C15ull);
    return (uint16_t)((h >> (16u * (uint32_t)(i & 3ull))) & mask);
}

// HiSeqX-shaped categorical column (KAT-E): exact category counts taken from
// the samtools output quoted in README.md:179-191, scattered by the bijection
// j = (i * M) mod N.
constexpr uint64_t kHiseqxN = 824541892ull;
constexpr uint64_t kHiseqxM = 509594915ull;

__host__ __device__ __forceinline__ uint16_t synth_hiseqx_at(uint64_t i, uint64_t seed,
                                                            uint32_t qcfail_ppm)
{
    const uint64_t j = ((i % kHiseqxN) * kHiseqxM) % kHiseqxN;
    // cumulative category boundaries
    uint16_t v;
    if (j < 781085884ull) {  // 4 x 195,271,471 proper pairs
        const uint32_t c = (uint32_t)(j / 195271471ull);
        v = c == 0 ? 99 : c == 1 ? 147 : c == 2 ? 83 : 163;
    } else if (j < 797950890ull) {  // 2 x 8,432,503 mapped, not proper
        v = (j - 781085884ull) < 8432503ull ? 97 : 145;
    } else if (j < 799989775ull) {  // singletons 1,019,443 + 1,019,442
        v = (j - 797950890ull) < 1019443ull ? 73 : 137;
    } else if (j < 802028660ull) {  // unmapped, mate mapped 1,019,443 + 1,019,442
        v = (j - 799989775ull) < 1019443ull ? 133 : 69;
    } else if (j < 819148264ull) {  // both unmapped 2 x 8,559,802
        v = (j - 802028660ull) < 8559802ull ? 77 : 141;
    } else {  // supplementary 4 x 1,348,407
        const uint32_t c = (uint32_t)((j - 819148264ull) / 1348407ull);
        v = c == 0 ? 2113 : c == 1 ? 2177 : c == 2 ? 2129 : 2193;
    }
    if (qcfail_ppm) {
        const uint64_t h = mix64(seed + (i + 1ull) * 0x9E3779B97F4A7C15ull);
        if ((uint32_t)(h % 1000000ull) < qcfail_ppm) v |= 0x200;
    }
    return v;
}

__global__ void synth_uniform_kernel(uint16_t* __restrict__ out, uint64_t start, uint64_t n,
                                     uint64_t seed, uint16_t mask)
{
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n;
         k += (uint64_t)gridDim.x * blockDim.x)
        out[k] = synth_uniform_at(seed, start + k, mask);
}

__global__ void synth_hiseqx_kernel(uint16_t* __restrict__ out, uint64_t start, uint64_t n,
                                    uint64_t seed, uint32_t qcfail_ppm)
{
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n;
         k += (uint64_t)gridDim.x * blockDim.x)
        out[k] = synth_hiseqx_at(start + k, seed, qcfail_ppm);
}

}  // namespace fsb200
