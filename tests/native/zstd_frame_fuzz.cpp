// AddressSanitizer / UBSan fuzz of the PRODUCT's Zstd frame decoder on the host
// (libflagstats_b200/csrc/zstd_frame.cuh): frames from the real libzstd, exact-size input and
// output buffers, byte flips and truncation.  Whatever the input, the decoder must stay inside
// its buffers and either reject the frame or produce at most `cap` bytes.  Built and run by
// tests/test_zstd_frame_host.py.  Test scaffolding; not part of the product library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "zstd_frame.cuh"

extern "C" {  // the image has libzstd.so.1 but no zstd.h
size_t ZSTD_compressBound(size_t n);
size_t ZSTD_compress(void* dst, size_t cap, const void* src, size_t n, int level);
unsigned ZSTD_isError(size_t code);
}

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "zstd_frame_fuzz: line %d: %s\n", __LINE__, #c); return 1; } } while (0)

static uint64_t rng_state = 0x2545F4914F6CDD1Dull;
static uint32_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 11);
}

int main()
{
    fsb200::zstd::Work* w = new fsb200::zstd::Work;
    for (int kind = 0; kind < 4; ++kind) {
        const uint64_t raw = kind == 3 ? 140001 : 400000;
        uint8_t* a = (uint8_t*)std::malloc(raw);
        CHECK(a);
        for (uint64_t i = 0; i < raw; ++i)
            a[i] = kind == 0 ? (uint8_t)((i / 2 % 7) * 16 + (i & 1) * 3)       // FLAG-like, periodic
                 : kind == 1 ? (uint8_t)(rnd() & 0x0F)                          // 4-bit noise: Huffman literals
                 : kind == 2 ? (uint8_t)((i / 97) % 5)                          // long runs
                             : (uint8_t)rnd();                                  // incompressible: raw blocks
        for (int level = 1; level <= 19; level += 6) {
            const size_t cap = ZSTD_compressBound(raw);
            uint8_t* tmp = (uint8_t*)std::malloc(cap);
            CHECK(tmp);
            const size_t c = ZSTD_compress(tmp, cap, a, raw, level);
            CHECK(!ZSTD_isError(c) && c > 8);
            uint8_t* frame = (uint8_t*)std::malloc(c);
            uint8_t* back = (uint8_t*)std::malloc(raw);
            CHECK(frame && back);
            std::memcpy(frame, tmp, c);
            CHECK(fsb200::zstd::decode_frame(frame, c, back, raw, *w) == (int64_t)raw);
            CHECK(std::memcmp(back, a, raw) == 0);
            for (int t = 0; t < 400; ++t) {
                uint8_t* f = (uint8_t*)std::malloc(c);
                CHECK(f);
                std::memcpy(f, frame, c);
                const int flips = 1 + (int)(rnd() % 4);
                for (int k = 0; k < flips; ++k) f[4 + rnd() % (c - 4)] ^= (uint8_t)(1u << (rnd() % 8));
                const uint64_t cut = (t % 4 == 0) ? (uint64_t)(rnd() % (c + 1)) : (uint64_t)c;
                const uint64_t ocap = (t % 5 == 0) ? (uint64_t)(rnd() % (raw + 1)) : raw;
                uint8_t* o = (uint8_t*)std::malloc(ocap ? ocap : 1);
                CHECK(o);
                const int64_t got = fsb200::zstd::decode_frame(f, cut, o, ocap, *w);
                CHECK(got <= (int64_t)ocap);
                {   // second version: exact-size literal buffer and descriptor array, same verdict, same bytes
                    using namespace fsb200::zstd;
                    const uint64_t lit_cap = ocap + kBlockMax + 64;
                    const uint32_t cap_d = (uint32_t)count_descriptors(f, cut, ocap);
                    uint8_t* lit = (uint8_t*)std::malloc(lit_cap);
                    SeqDesc* d = (SeqDesc*)std::malloc(sizeof(SeqDesc) * cap_d);
                    uint8_t* o2 = (uint8_t*)std::malloc(ocap ? ocap : 1);
                    CHECK(lit && d && o2);
                    uint32_t nd = 0;
                    uint64_t lit_used = 0;
                    int64_t got2 = parse_frame(f, cut, ocap, w->t, lit, lit_cap, d, cap_d, &nd, &lit_used);
                    CHECK(nd <= cap_d && lit_used <= lit_cap);
                    if (got2 >= 0) got2 = apply_descriptors(d, nd, lit, lit_used, o2, (uint64_t)got2);
                    CHECK((got2 < 0) == (got < 0));
                    if (got >= 0) CHECK(got2 == got && std::memcmp(o, o2, (size_t)got) == 0);
                    std::free(o2);
                    std::free(d);
                    std::free(lit);
                }
                std::free(o);
                std::free(f);
            }
            std::free(back);
            std::free(frame);
            std::free(tmp);
        }
        std::free(a);
    }
    delete w;
    std::puts("zstd_frame_fuzz ok");
    return 0;
}
