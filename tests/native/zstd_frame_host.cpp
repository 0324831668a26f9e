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
