// Host build of the PRODUCT's Zstd frame decoder (libflagstats_b200/csrc/zstd_frame.cuh is
// __host__ __device__): compiled with g++ by tests/test_zstd_frame_host.py and held to the real
// libzstd on the CPU, so that the GPU tests only have to show the same code gives the same
// bytes on the device.  Test scaffolding; not part of the product library.
#include <cstdint>
#include <new>

#include "zstd_frame.cuh"

extern "C" int64_t zstd_frame_host(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap)
{
    fsb200::zstd::Work* w = new (std::nothrow) fsb200::zstd::Work;
    if (!w) return -100;
    const int64_t r = fsb200::zstd::decode_frame(in, n, out, cap, *w);
    delete w;
    return r;
}

extern "C" uint64_t zstd_frame_work_bytes(void) { return sizeof(fsb200::zstd::Work); }

// Second version: entropy stage into descriptors + literal buffer (what zstd_parse_kernel runs, one lane per
// frame), then the descriptors applied sequentially (the device applies them with l4_copy).
extern "C" int64_t zstd_frame_host_v2(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap)
{
    using namespace fsb200::zstd;
    Tables* t = new (std::nothrow) Tables;
    const uint64_t lit_cap = cap + kBlockMax + 64;
    // exactly the slice the library gives the frame: the count from the headers
    const uint32_t cap_d = (uint32_t)count_descriptors(in, n, cap);
    uint8_t* lit = new (std::nothrow) uint8_t[lit_cap];
    SeqDesc* d = new (std::nothrow) SeqDesc[cap_d];
    int64_t r = -100;
    if (t && lit && d) {
        uint32_t nd = 0;
        uint64_t lit_used = 0;
        r = parse_frame(in, n, cap, *t, lit, lit_cap, d, cap_d, &nd, &lit_used);
        if (r >= 0) {
            // the descriptors must tile [0, r) in order
            uint64_t pos = 0;
            for (uint32_t k = 0; k < nd && r >= 0; ++k) {
                if (d[k].out_pos != pos && !(d[k].out_pos > pos)) r = -101;  // positions never go back
                pos = d[k].out_pos;
            }
            if (r >= 0) r = apply_descriptors(d, nd, lit, lit_used, out, (uint64_t)r);
        }
    }
    delete t;
    delete[] lit;
    delete[] d;
    return r;
}

extern "C" uint64_t zstd_frame_tables_bytes(void) { return sizeof(fsb200::zstd::Tables); }
