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

int main(int argc, char** argv)
{
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
