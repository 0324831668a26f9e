// dropin_check.cpp -- proves the drop-in: this program is compiled against the
// REFERENCE's libflagstats.h with integration/libflagstats_h_cuda.patch applied
// (three added hunks, nothing removed) and -DFLAGSTATS_HAVE_CUDA, and linked
// with libflagstats_cuda.so.  It then uses only the reference's own entry
// points -- FLAGSTATS_get_function and FLAGSTATS_u16 (libflagstats.h:2976,3024)
// -- the way the reference's drivers do (benchmark/flagstats.cpp:304,328-329:
// one shared counters[32] accumulated over 1,024,000-byte blocks) and checks
// the result against FLAGSTAT_scalar.
//
// With a GPU present the dispatcher must hand out FLAGSTAT_cuda for long
// blocks; without one it must fall through to the reference's CPU kernels.
// Built by integration/build_dropin.sh into oracle/_ref/dropin_check.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "libalgebra.h"
#include "libflagstats.h"

static const int kCore19[] = {2, 6, 7, 8, 10, 11, 12, 13, 14, 18, 22, 23, 24, 25, 26, 27, 28, 29, 30};

static const char* name_of(FLAGSTATS_func f)
{
    if (f == &FLAGSTAT_scalar) return "FLAGSTAT_scalar";
#if defined(FLAGSTATS_HAVE_CUDA)
    if (f == &FLAGSTAT_cuda) return "FLAGSTAT_cuda";
#endif
#if defined(STORM_HAVE_SSE42)
    if (f == &FLAGSTAT_sse4) return "FLAGSTAT_sse4";
#endif
#if defined(STORM_HAVE_AVX2)
    if (f == &FLAGSTAT_avx2) return "FLAGSTAT_avx2";
#endif
#if defined(STORM_HAVE_AVX512)
    if (f == &FLAGSTAT_avx512) return "FLAGSTAT_avx512";
#endif
    return "?";
}

static int compare(const uint32_t* want_scalar, const uint32_t* got, uint32_t n, const char* what)
{
    int bad = 0;
    for (int i : kCore19)
        if (want_scalar[i] != got[i]) {
            std::printf("MISMATCH %s slot %d: scalar %u got %u\n", what, i, want_scalar[i], got[i]);
            ++bad;
        }
    if (n >= 256 && got[9] != n - want_scalar[25]) {  // SIMD / CUDA convention, libflagstats.h:429
        std::printf("MISMATCH %s slot 9: want %u got %u\n", what, n - want_scalar[25], got[9]);
        ++bad;
    }
    return bad;
}

// ---- --time: throughput of the drop-in block loop and the GPU / CPU crossover ----------
// benchmark/flagstats.cpp:304,328-329 calls `FLAGSTATS_get_function(N)` once per 512,000-record
// block on ordinary (pageable) host memory and waits for the counters.  Timed here: that
// loop through the patched dispatcher (FLAGSTAT_cuda), the same loop with the device branch
// switched off (FLAGSTAT_cuda_set_min_len(UINT32_MAX): the reference's own best CPU kernel, 1
// thread, as shipped), the same blocks from pinned memory, and one call per n for n = 2^10 ..
// 2^24 to find the length where the synchronous GPU call starts to win.  JSON lines.
#include <algorithm>
#include <chrono>

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static double time_block_loop(const uint16_t* data, uint32_t n_blocks, uint32_t block, uint32_t* counters)
{
    const double t0 = now_s();
    for (uint32_t b = 0; b < n_blocks; ++b) {
        FLAGSTATS_func func = FLAGSTATS_get_function(block);
        (*func)(data + (size_t)b * block, block, counters);
    }
    return now_s() - t0;
}

static double median_call_us(const uint16_t* data, uint32_t n, int reps, size_t span)
{
    // successive calls walk through `span` records so that neither side sees a warm cache only
    std::vector<double> t;
    uint32_t c[32] = {0};
    size_t off = 0;
    for (int r = 0; r < reps + 2; ++r) {
        if (off + n > span) off = 0;
        const double t0 = now_s();
        FLAGSTATS_func func = FLAGSTATS_get_function(n);
        (*func)(data + off, n, c);
        const double dt = now_s() - t0;
        if (r >= 2) t.push_back(dt * 1e6);
        off += n;
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

static int time_mode()
{
    const uint32_t block = 512000, n_blocks = 256;
    const size_t total = (size_t)block * n_blocks;
    std::vector<uint16_t> flags(total);
    uint64_t x = 88172645463325252ull;
    for (auto& v : flags) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        v = (uint16_t)(x & 0x0FFF);
    }
    const uint32_t thr0 = FLAGSTAT_cuda_min_len();
    const int have_gpu = FLAGSTAT_cuda_available();
    std::printf("{\"devices\": %d, \"cuda_min_len_default\": %u, \"block_records\": %u, \"blocks\": %u}\n",
                have_gpu, thr0, block, n_blocks);
    uint32_t want[32] = {0};
    // CPU: the dispatcher with the device branch off
    FLAGSTAT_cuda_set_min_len(0xFFFFFFFFu);
    const char* cpu_kernel = name_of(FLAGSTATS_get_function(block));
    time_block_loop(flags.data(), 16, block, want);
    std::memset(want, 0, sizeof want);
    double cpu_s = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        uint32_t c[32] = {0};
        cpu_s = std::min(cpu_s, time_block_loop(flags.data(), n_blocks, block, c));
        std::memcpy(want, c, sizeof c);
    }
    std::printf("{\"loop\": \"cpu\", \"kernel\": \"%s\", \"threads\": 1, \"us_per_block\": %.2f, \"gbs\": %.2f, \"grec_s\": %.3f}\n",
                cpu_kernel, cpu_s / n_blocks * 1e6, total * 2.0 / cpu_s / 1e9, total / cpu_s / 1e9);
    if (!have_gpu) return 0;
    // GPU: pageable blocks through the patched dispatcher
    FLAGSTAT_cuda_set_min_len(1);
    int bad = 0;
    {
        uint32_t c[32] = {0};
        time_block_loop(flags.data(), 16, block, c);
        double s_best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            std::memset(c, 0, sizeof c);
            s_best = std::min(s_best, time_block_loop(flags.data(), n_blocks, block, c));
        }
        for (int i : kCore19) bad += c[i] != want[i];
        std::printf("{\"loop\": \"cuda_pageable\", \"kernel\": \"%s\", \"us_per_block\": %.2f, \"gbs\": %.2f, \"grec_s\": %.3f, "
                    "\"same_counters\": %s}\n", name_of(FLAGSTATS_get_function(block)), s_best / n_blocks * 1e6,
                    total * 2.0 / s_best / 1e9, total / s_best / 1e9, bad ? "false" : "true");
    }
    // GPU: the same blocks from pinned memory (no driver staging copy)
    if (uint16_t* pin = (uint16_t*)FLAGSTAT_cuda_malloc_host(total * 2)) {
        std::memcpy(pin, flags.data(), total * 2);
        uint32_t c[32] = {0};
        time_block_loop(pin, 16, block, c);
        double s_best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
            std::memset(c, 0, sizeof c);
            s_best = std::min(s_best, time_block_loop(pin, n_blocks, block, c));
        }
        int b2 = 0;
        for (int i : kCore19) b2 += c[i] != want[i];
        bad += b2;
        std::printf("{\"loop\": \"cuda_pinned\", \"us_per_block\": %.2f, \"gbs\": %.2f, \"grec_s\": %.3f, \"same_counters\": %s}\n",
                    s_best / n_blocks * 1e6, total * 2.0 / s_best / 1e9, total / s_best / 1e9, b2 ? "false" : "true");
        // per-call time by length, pinned
        for (uint32_t n = 1u << 10; n <= (1u << 24); n <<= 1) {
            FLAGSTAT_cuda_set_min_len(1);
            const double g = median_call_us(pin, n, 21, total);
            std::printf("{\"sweep\": \"pinned\", \"n\": %u, \"cuda_us\": %.2f}\n", n, g);
        }
        FLAGSTAT_cuda_free_host(pin);
    }
    // per-call time by length, pageable: device vs the reference's CPU kernel
    uint32_t crossover = 0;
    for (uint32_t n = 1u << 10; n <= (1u << 24); n <<= 1) {
        FLAGSTAT_cuda_set_min_len(1);
        const double g = median_call_us(flags.data(), n, 21, total);
        FLAGSTAT_cuda_set_min_len(0xFFFFFFFFu);
        const double c = median_call_us(flags.data(), n, 21, total);
        if (!crossover && g < c) crossover = n;
        if (crossover && g >= c) crossover = 0;  // must stay ahead from there on
        std::printf("{\"sweep\": \"pageable\", \"n\": %u, \"cuda_us\": %.2f, \"cpu_us\": %.2f, \"cpu_kernel\": \"%s\"}\n", n, g, c,
                    name_of(FLAGSTATS_get_function(n)));
    }
    std::printf("{\"crossover_records\": %u, \"note\": \"smallest power of two from which the synchronous FLAGSTAT_cuda "
                "call on pageable memory stays faster than the reference's 1-thread CPU kernel\"}\n", crossover);
    FLAGSTAT_cuda_set_min_len(thr0);
    return bad ? 1 : 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::strcmp(argv[1], "--time") == 0) return time_mode();
    const int have_gpu = FLAGSTAT_cuda_available();
    const uint32_t thr = FLAGSTAT_cuda_min_len();
    std::printf("devices=%d cuda_min_len=%u\n", have_gpu, thr);

    const uint32_t total = 5300123;
    std::vector<uint16_t> flags(total + 8);
    uint64_t x = 88172645463325252ull;
    for (auto& v : flags) {  // xorshift64, U(0,4095) like benchmark/generate.cpp:11
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        v = (uint16_t)(x & 0x0FFF);
    }

    int bad = 0, cuda_selected = 0;
    const uint32_t lens[] = {0, 100, 255, 256, 1024, 4097, 65535, 65536, 512000, 1000003, total};
    for (uint32_t n : lens) {
        for (int off = 0; off < 2; ++off) {
            const uint16_t* p = flags.data() + off;  // +1 element: 2-byte aligned only
            uint32_t want[32] = {0}, got[32] = {0}, got2[32] = {0};
            FLAGSTAT_scalar(p, n, want);
            FLAGSTATS_func f = FLAGSTATS_get_function(n);
            (*f)(p, n, got);
            FLAGSTATS_u16(p, n, got2);
            if (f == &FLAGSTAT_cuda) ++cuda_selected;
            if (off == 0) std::printf("n=%-8u -> %s\n", n, name_of(f));
            bad += compare(want, got, f == &FLAGSTAT_scalar ? 0 : n, "get_function");
            bad += compare(want, got2, f == &FLAGSTAT_scalar ? 0 : n, "FLAGSTATS_u16");
        }
    }

    // the block loop of benchmark/flagstats.cpp:304-329
    {
        uint32_t counters[32] = {0}, want[32] = {0};
        const uint32_t block = 512000;
        for (uint32_t lo = 0; lo < total; lo += block) {
            const uint32_t N = (total - lo < block) ? (total - lo) : block;
            FLAGSTATS_func func = FLAGSTATS_get_function(N);
            (*func)(flags.data() + lo, N, counters);
        }
        FLAGSTAT_scalar(flags.data(), total, want);
        bad += compare(want, counters, total, "block loop");
    }

    // the same block loop with the threshold lowered the way a caller with short blocks would
    // (FLAGSTAT_cuda_set_min_len): every 512,000-record block must now go to the device
    if (have_gpu) {
        FLAGSTAT_cuda_set_min_len(4096);
        uint32_t counters[32] = {0}, want[32] = {0};
        const uint32_t block = 512000;
        int on_device = 0;
        for (uint32_t lo = 0; lo < total; lo += block) {
            const uint32_t N = (total - lo < block) ? (total - lo) : block;
            FLAGSTATS_func func = FLAGSTATS_get_function(N);
            on_device += func == &FLAGSTAT_cuda;
            (*func)(flags.data() + lo, N, counters);
        }
        FLAGSTAT_scalar(flags.data(), total, want);
        bad += compare(want, counters, total, "block loop on the device");
        if (on_device != (int)((total + block - 1) / block)) {
            std::printf("FAIL: min_len = 4096 but only %d blocks went to FLAGSTAT_cuda\n", on_device);
            ++bad;
        }
        cuda_selected += on_device;
        FLAGSTAT_cuda_set_min_len(thr);
    }

    if (have_gpu && cuda_selected == 0) {
        std::printf("FAIL: a device is present but FLAGSTAT_cuda was never selected\n");
        ++bad;
    }
    if (!have_gpu && cuda_selected != 0) {
        std::printf("FAIL: no device but FLAGSTAT_cuda was selected\n");
        ++bad;
    }
    std::printf("%s (cuda selected for %d calls)\n", bad ? "FAIL" : "OK", cuda_selected);
    (void)argc; (void)argv;
    return bad ? 1 : 0;
}
