/*
 * sanitize_check.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The oracle's two C files compiled as ONE translation unit with
 * -fsanitize=address,undefined and driven over exact-size heap buffers (so that any
 * out-of-bounds access, misaligned load or signed overflow in the checker itself is
 * caught): the flagstat restatements on ragged / odd-based arrays, the samtools loop,
 * pospopcnt, both generators, LZ4 round trips and a byte-flip fuzz of the LZ4 decoder and
 * the container walk (malformed input must be rejected, never read or written out of
 * bounds).  Built and run by tests/test_oracle_sanitizers.py.
 */
#include <stdio.h>
#include <stdlib.h>

#include "flagstat_oracle.c"
#include "lz4_oracle.c"
#include "zstd_oracle.c"

#ifdef HAVE_LIBZSTD  /* the image has the runtime but no zstd.h: the three stable prototypes, by hand */
size_t ZSTD_compressBound(size_t n);
size_t ZSTD_compress(void* dst, size_t cap, const void* src, size_t n, int level);
unsigned ZSTD_isError(size_t code);
#endif

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "sanitize_check: %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void)
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 11);
}

static void count_block(const uint16_t* a, uint64_t n, void* ctx)
{
    oracle_flagstat_simd_u64(a, n, (uint64_t*)ctx);
}

int main(void)
{
    static const uint64_t lens[] = {0, 1, 2, 7, 8, 9, 255, 256, 257, 1023, 4096, 65537, 300001};
    for (unsigned li = 0; li < sizeof lens / sizeof lens[0]; ++li) {
        const uint64_t n = lens[li];
        for (int kind = 0; kind < 3; ++kind) {
            /* exact-size allocation: one element past the end is a red zone */
            uint16_t* a = (uint16_t*)malloc((n ? n : 1) * sizeof(uint16_t));
            CHECK(a);
            if (kind == 0) oracle_synth_uniform(a, 12345 + li, n, 7, 0x0FFF);
            else if (kind == 1) oracle_synth_uniform(a, 0, n, li, 0xFFFF);
            else oracle_synth_hiseqx(a, 824541892ull - n / 2, n, 3, 20000);  /* crosses the period */
            uint64_t s[32] = {0}, v[32] = {0}, m[32] = {0}, pp[16];
            uint32_t s32[32] = {0}, pp32[16];
            long long st[26] = {0};
            CHECK(oracle_flagstat_scalar_u64(a, n, s) == 0);
            CHECK(oracle_flagstat_simd_u64(a, n, v) == 0);
            CHECK(oracle_flagstat_maskselect_u64(a, n, m) == 0);
            CHECK(oracle_flagstat_simd_u32(a, (uint32_t)n, s32) == 0);
            CHECK(oracle_samtools_loop(a, n, st) == 0);
            CHECK(oracle_pospopcnt_u16_u64(a, n, pp) == 0);
            CHECK(oracle_pospopcnt_u16(a, n, pp32) == 0);
            for (int i = 0; i < 32; ++i) {
                CHECK(v[i] == m[i]);
                CHECK((uint32_t)v[i] == s32[i]);
                if (i != 9) CHECK(s[i] == v[i]);
            }
            CHECK(v[9] + v[25] == n);
            CHECK((uint64_t)(st[0] + st[1]) == n);                  /* n_reads */
            CHECK((uint64_t)st[22] == v[8] && (uint64_t)st[23] == v[24]);  /* n_secondary */
            for (int j = 0; j < 16; ++j) CHECK(pp[j] == pp32[j] && pp[j] <= n);

            /* LZ4: round trip through the restated compressor and decoder */
            const uint64_t raw = n * 2;
            const uint64_t cap = raw + raw / 255 + 64;
            uint8_t* comp = (uint8_t*)malloc(cap);
            uint8_t* back = (uint8_t*)malloc(raw ? raw : 1);
            CHECK(comp && back);
            const int64_t c = oracle_lz4_compress((const uint8_t*)a, raw, comp, cap);
            CHECK(c > 0 || raw == 0);
            if (c > 0) {
                uint8_t* exact = (uint8_t*)malloc((size_t)c);  /* decoder input with a red zone right behind it */
                CHECK(exact);
                memcpy(exact, comp, (size_t)c);
                CHECK(oracle_lz4_decompress(exact, (uint64_t)c, back, raw) == (int64_t)raw);
                CHECK(memcmp(back, a, raw) == 0);
                /* container walk over [raw][comp][block] x 2 */
                const uint64_t cn = 2 * (8 + (uint64_t)c);
                uint8_t* cont = (uint8_t*)malloc(cn);
                CHECK(cont);
                for (int rep = 0; rep < 2; ++rep) {
                    const int32_t hr = (int32_t)raw, hc = (int32_t)c;
                    memcpy(cont + rep * (8 + c), &hr, 4);
                    memcpy(cont + rep * (8 + c) + 4, &hc, 4);
                    memcpy(cont + rep * (8 + c) + 8, exact, (size_t)c);
                }
                uint64_t w[32] = {0};
                CHECK(oracle_lz4_container_walk(cont, cn, back, raw, count_block, w) == (int64_t)(2 * n));
                for (int i = 0; i < 32; ++i) CHECK(w[i] == 2 * v[i]);
                /* fuzz: flipped bytes / truncation must be rejected or decode to <= raw bytes, never overrun */
                for (int t = 0; t < 300 && n <= 65537; ++t) {
                    uint8_t* f = (uint8_t*)malloc((size_t)c);
                    CHECK(f);
                    memcpy(f, exact, (size_t)c);
                    const int flips = 1 + (int)(rnd() % 4);
                    for (int k = 0; k < flips; ++k) f[rnd() % (uint64_t)c] ^= (uint8_t)(1u << (rnd() % 8));
                    const uint64_t cut = (t % 3 == 0) ? (uint64_t)(rnd() % ((uint64_t)c + 1)) : (uint64_t)c;
                    const int64_t got = oracle_lz4_decompress(f, cut, back, raw);
                    CHECK(got <= (int64_t)raw);
                    memcpy(cont + 8, f, (size_t)c);
                    (void)oracle_lz4_container_walk(cont, cn - (uint64_t)(rnd() % 9), back, raw, NULL, NULL);
                    memcpy(cont + 8, exact, (size_t)c);
                    free(f);
                }
                free(cont);
                free(exact);
            }
            free(comp);
            free(back);
            free(a);
        }
    }
#ifdef HAVE_LIBZSTD
    /* Zstd frames from the real compressor: decode, then byte-flip / truncation fuzz */
    for (int kind = 0; kind < 4; ++kind) {
        const uint64_t n = kind == 3 ? 70001 : 200000;
        uint16_t* a = (uint16_t*)malloc(n * sizeof(uint16_t));
        CHECK(a);
        if (kind == 0) oracle_synth_hiseqx(a, 0, n, 1, 3000);
        else if (kind == 1) oracle_synth_uniform(a, 0, n, 5, 0x0FFF);
        else if (kind == 2) for (uint64_t i = 0; i < n; ++i) a[i] = (uint16_t)((i / 37) % 11 * 16 + 99);
        else oracle_synth_uniform(a, 0, n, 6, 0x0003);
        const uint64_t raw = n * 2;
        for (int level = 1; level <= 19; level += 6) {
            const size_t cap = ZSTD_compressBound(raw);
            uint8_t* tmp = (uint8_t*)malloc(cap);
            CHECK(tmp);
            const size_t c = ZSTD_compress(tmp, cap, a, raw, level);
            CHECK(!ZSTD_isError(c) && c > 0);
            uint8_t* frame = (uint8_t*)malloc(c);  /* exact size: red zone right behind the frame */
            uint8_t* back = (uint8_t*)malloc(raw);
            CHECK(frame && back);
            memcpy(frame, tmp, c);
            CHECK(oracle_zstd_decompress(frame, c, back, raw) == (int64_t)raw);
            CHECK(memcmp(back, a, raw) == 0);
            for (int t = 0; t < 400; ++t) {
                uint8_t* f = (uint8_t*)malloc(c);
                CHECK(f);
                memcpy(f, frame, c);
                const int flips = 1 + (int)(rnd() % 4);
                for (int k = 0; k < flips; ++k) f[4 + rnd() % (c - 4)] ^= (uint8_t)(1u << (rnd() % 8));
                const uint64_t cut = (t % 4 == 0) ? (uint64_t)(rnd() % (c + 1)) : (uint64_t)c;
                const uint64_t ocap = (t % 5 == 0) ? (uint64_t)(rnd() % (raw + 1)) : raw;
                uint8_t* o = (uint8_t*)malloc(ocap ? ocap : 1);  /* exact-size output too */
                CHECK(o);
                const int64_t got = oracle_zstd_decompress(f, cut, o, ocap);
                CHECK(got <= (int64_t)ocap);
                free(o);
                free(f);
            }
            free(back);
            free(frame);
            free(tmp);
        }
        free(a);
    }
#endif
    /* every 16-bit word through the per-record forms */
    for (uint32_t x = 0; x < 65536; ++x) {
        uint64_t f[32] = {0};
        oracle_flagstat_update((uint16_t)x, f);
        const uint16_t y = oracle_mask_select((uint16_t)x);
        CHECK((y & 0x8000u) == 0);
    }
    puts("sanitize_check ok");
    return 0;
}
