/*
 * ref_shim.cpp -- exposes the UNMODIFIED reference (mklarqvist/libflagstats)
 * through a C ABI so Python tests and bench.py can call it.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  This file contains no flagstat logic of
 * its own: it #includes the reference's two headers where they lie under
 * /root/reference (-I flags in oracle/Makefile) and forwards to them.  The
 * build product goes to oracle/_ref/ (git-ignored, shipped to the GPU box by
 * gpurun).  The reference's kernels are `static` functions in a header inside
 * extern "C" and FLAGSTAT_scalar_update is a non-inline definition
 * (libflagstats.h:118), so everything has to live in this single TU.
 *
 * The reference selects its ISA per function with
 * __attribute__((target(...))) (libflagstats.h:182,965,1644) and dispatches on
 * cpuid at run time (libflagstats.h:2976-3022), so this TU is portable across
 * x86-64 hosts whatever -march it is compiled with.
 *
 * The only code added here is a range-sharded pthread wrapper (ref_flagstat_mt):
 * the reference itself is single-threaded and ships no threading (SURVEY.md
 * section 2.3); the wrapper gives every thread a private flags[32], runs the
 * unmodified kernel on a contiguous range and sums the counters.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <pthread.h>

#include "libalgebra.h"    /* /root/reference/libalgebra/libalgebra.h */
#include "libflagstats.h"  /* /root/reference/libflagstats.h */

/* The samtools-style caller of the reference benchmark: bam_flagstat_t, the
 * flagstat_loop macro and percent() (benchmark/flagstats.cpp:43-71,73-78).  That
 * file cannot be compiled here (it needs lz4.h / zstd.h), so oracle/Makefile cuts
 * exactly those definitions out of it at build time into a temporary include,
 * compiles them into this TU unmodified and deletes the temporary again. */
#ifdef REF_SAMTOOLS_INC
#include <cstdio>
#include REF_SAMTOOLS_INC
#endif

namespace {

struct Entry {
    const char* name;
    FLAGSTATS_func fn;
    int need;  /* cpuid bit required, 0 = none */
};

const Entry kTable[] = {
    {"scalar", &FLAGSTAT_scalar, 0},
#if defined(STORM_HAVE_SSE42)
    {"sse4", &FLAGSTAT_sse4, STORM_CPUID_runtime_bit_SSE42},
    {"sse4_improved", &FLAGSTAT_sse4_improved, STORM_CPUID_runtime_bit_SSE42},
    {"sse4_improved2", &FLAGSTAT_sse4_improved2, STORM_CPUID_runtime_bit_SSE42},
#endif
#if defined(STORM_HAVE_AVX2)
    {"avx2", &FLAGSTAT_avx2, STORM_CPUID_runtime_bit_AVX2},
    {"avx2_improved", &FLAGSTAT_avx2_improved, STORM_CPUID_runtime_bit_AVX2},
    {"avx2_improved2", &FLAGSTAT_avx2_improved2, STORM_CPUID_runtime_bit_AVX2},
#endif
#if defined(STORM_HAVE_AVX512)
    {"avx512", &FLAGSTAT_avx512, STORM_CPUID_runtime_bit_AVX512BW},
    {"avx512_improved", &FLAGSTAT_avx512_improved, STORM_CPUID_runtime_bit_AVX512BW},
    {"avx512_improved2", &FLAGSTAT_avx512_improved2, STORM_CPUID_runtime_bit_AVX512BW},
    {"avx512_improved3", &FLAGSTAT_avx512_improved3, STORM_CPUID_runtime_bit_AVX512BW},
    {"avx512_improved4", &FLAGSTAT_avx512_improved4, STORM_CPUID_runtime_bit_AVX512BW},
#endif
};
const int kTableLen = (int)(sizeof(kTable) / sizeof(kTable[0]));

int cpuid_once()
{
    static const int c = STORM_get_cpuid();
    return c;
}

const Entry* find(const char* name)
{
    for (int i = 0; i < kTableLen; ++i)
        if (std::strcmp(kTable[i].name, name) == 0) return &kTable[i];
    return nullptr;
}

bool runnable(const Entry* e)
{
    return e && (e->need == 0 || (cpuid_once() & e->need) == e->need);
}

struct Job {
    FLAGSTATS_func fn;
    const uint16_t* base;
    uint64_t len;
    uint64_t flags[32];
};

/* uint32_t len / uint32_t counters: feed the reference at most 2^30 records per
 * call and widen between calls. */
void run_range(Job* j)
{
    std::memset(j->flags, 0, sizeof j->flags);
    const uint64_t kChunk = 1ull << 30;
    for (uint64_t off = 0; off < j->len; off += kChunk) {
        const uint32_t n = (uint32_t)((j->len - off < kChunk) ? (j->len - off) : kChunk);
        uint32_t f[32];
        std::memset(f, 0, sizeof f);
        j->fn(j->base + off, n, f);
        for (int k = 0; k < 32; ++k) j->flags[k] += f[k];
    }
}

void* thread_main(void* p)
{
    run_range((Job*)p);
    return nullptr;
}

}  // namespace

extern "C" {

int ref_cpuid(void) { return cpuid_once(); }
int ref_has_sse42(void) { return (cpuid_once() & STORM_CPUID_runtime_bit_SSE42) != 0; }
int ref_has_avx2(void) { return (cpuid_once() & STORM_CPUID_runtime_bit_AVX2) != 0; }
int ref_has_avx512bw(void) { return (cpuid_once() & STORM_CPUID_runtime_bit_AVX512BW) != 0; }

int ref_num_kernels(void) { return kTableLen; }
const char* ref_kernel_name(int i) { return (i >= 0 && i < kTableLen) ? kTable[i].name : ""; }
int ref_kernel_runnable(const char* name) { return runnable(find(name)) ? 1 : 0; }

/* Call one named reference kernel.  Returns -1 if unknown, -2 if this CPU
 * cannot run it, else the kernel's own return value (always 0). */
int ref_flagstat(const char* name, const uint16_t* array, uint32_t len, uint32_t* flags)
{
    const Entry* e = find(name);
    if (!e) return -1;
    if (!runnable(e)) return -2;
    return e->fn(array, len, flags);
}

/* FLAGSTATS_u16, libflagstats.h:3024-3070 */
uint64_t ref_FLAGSTATS_u16(const uint16_t* array, uint32_t n_len, uint32_t* flags)
{
    return FLAGSTATS_u16(array, n_len, flags);
}

/* Name of the kernel FLAGSTATS_get_function(n) returns, libflagstats.h:2976-3022 */
const char* ref_dispatch_name(uint32_t n_len)
{
    const FLAGSTATS_func f = FLAGSTATS_get_function(n_len);
    for (int i = 0; i < kTableLen; ++i)
        if (kTable[i].fn == f) return kTable[i].name;
    return "?";
}

/* STORM_pospopcnt_u16, libalgebra.h:3496-3551 */
int ref_pospopcnt_u16(const uint16_t* data, size_t len, uint32_t* out)
{
    return STORM_pospopcnt_u16(data, len, out);
}

/*
 * Range-sharded multi-thread wrapper around one unmodified reference kernel.
 * flags64 is ACCUMULATED into (like the kernels do).  *seconds receives the
 * wall time of the parallel section (thread create -> last join).
 */
int ref_flagstat_mt(const char* name, const uint16_t* array, uint64_t len,
                    int nthreads, uint64_t* flags64, double* seconds)
{
    const Entry* e = find(name);
    if (!e) return -1;
    if (!runnable(e)) return -2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;

    Job* jobs = (Job*)std::calloc((size_t)nthreads, sizeof(Job));
    pthread_t* th = (pthread_t*)std::calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { std::free(jobs); std::free(th); return -3; }

    /* contiguous ranges, boundaries rounded to 512 records (one AVX-512 block) */
    uint64_t per = (len / (uint64_t)nthreads) & ~511ull;
    uint64_t off = 0;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].fn = e->fn;
        jobs[t].base = array + off;
        jobs[t].len = (t == nthreads - 1) ? (len - off) : per;
        off += jobs[t].len;
    }

    const auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        run_range(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], nullptr, thread_main, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    for (int t = 0; t < nthreads; ++t)
        for (int k = 0; k < 32; ++k) flags64[k] += jobs[t].flags[k];
    std::free(jobs);
    std::free(th);
    return 0;
}

/* flagstat_loop over a column into a bam_flagstat_t seen as 26 long long
 * (benchmark/flagstats.cpp:43-71); accumulates.  -1 when the shim was built
 * without the benchmark source. */
int ref_samtools_loop(const uint16_t* array, uint64_t len, long long* s26)
{
#ifdef REF_SAMTOOLS_INC
    static_assert(sizeof(bam_flagstat_t) == 26 * sizeof(long long), "bam_flagstat_t layout");
    bam_flagstat_t* s = reinterpret_cast<bam_flagstat_t*>(s26);
    for (uint64_t i = 0; i < len; ++i) flagstat_loop(s, array[i]);
    return 0;
#else
    (void)array; (void)len; (void)s26;
    return -1;
#endif
}

/* percent(), benchmark/flagstats.cpp:73-78; buf >= 16 bytes */
int ref_samtools_percent(long long n, long long total, char* buf)
{
#ifdef REF_SAMTOOLS_INC
    percent(buf, n, total);
    return 0;
#else
    (void)n; (void)total; (void)buf;
    return -1;
#endif
}

}  /* extern "C" */
