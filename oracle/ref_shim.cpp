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
 * The only code added here are pthread wrappers: the reference itself is
 * single-threaded and ships no threading.
 *   ref_flagstat_mt / ref_pospopcnt_mt  every thread gets a private flags[32] /
 *       out[16], runs the unmodified kernel on a contiguous range; the counters
 *       are summed (1 thread and all host cores);
 *   ref_container_mt  the block loop of the reference's LZ4 / Zstd readers
 *       (benchmark/flagstats.cpp:311-331, 655-669: decompress with the system
 *       codec, then FLAGSTATS_get_function(N) per block) over a container in
 *       memory, as shipped (1 thread) or with the blocks dealt to pthreads.
 */
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <pthread.h>
#include <dlfcn.h>

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

/* ---- all-core forms of the other two baselines BASELINE.md section 4 names ------------- */

namespace {

struct PopJob {
    const uint16_t* base;
    uint64_t len;
    uint64_t out[16];
};

void* pop_thread(void* p)
{
    PopJob* j = (PopJob*)p;
    std::memset(j->out, 0, sizeof j->out);
    const uint64_t kChunk = 1ull << 30; /* uint32_t counters: widen between calls */
    for (uint64_t off = 0; off < j->len; off += kChunk) {
        const size_t n = (size_t)((j->len - off < kChunk) ? (j->len - off) : kChunk);
        uint32_t o[16];
        STORM_pospopcnt_u16(j->base + off, n, o); /* zeroes o first, libalgebra.h:3498 */
        for (int k = 0; k < 16; ++k) j->out[k] += o[k];
    }
    return nullptr;
}

/* The codecs the reference benchmark links (benchmark/flagstats.cpp:16-18) are system
 * libraries; their headers are not in this image but the runtime libraries are, so the two
 * entry points the reference's readers call are resolved with dlopen. */
typedef int (*lz4_safe_fn)(const char*, char*, int, int);
typedef size_t (*zstd_dec_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*zstd_err_fn)(size_t);

struct Codecs {
    lz4_safe_fn lz4 = nullptr;
    zstd_dec_fn zstd = nullptr;
    zstd_err_fn zstd_is_error = nullptr;
};

const Codecs& codecs()
{
    static const Codecs c = [] {
        Codecs k;
        if (void* h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL))
            k.lz4 = (lz4_safe_fn)dlsym(h, "LZ4_decompress_safe");
        if (void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL)) {
            k.zstd = (zstd_dec_fn)dlsym(h, "ZSTD_decompress");
            k.zstd_is_error = (zstd_err_fn)dlsym(h, "ZSTD_isError");
        }
        return k;
    }();
    return c;
}

struct Blk {
    const unsigned char* payload;
    int32_t raw, comp;
};

struct ContJob {
    const Blk* blk;
    size_t first, step, count;
    int codec;
    uint64_t flags[32];
    double decode_s;
    int rc;
};

/* The body of lz4_decompress() / zstd_decompress() (benchmark/flagstats.cpp:311-331,
 * 655-669) for the blocks first, first+step, ...: decompress into out_buffer, N = raw >> 1,
 * func = FLAGSTATS_get_function(N); (*func)(out_buffer, N, counters). */
void* cont_thread(void* p)
{
    ContJob* j = (ContJob*)p;
    std::memset(j->flags, 0, sizeof j->flags);
    j->decode_s = 0.0;
    j->rc = 0;
    uint8_t* out_buffer = (uint8_t*)STORM_aligned_malloc(STORM_get_alignment(), 1024000 + 65536);
    if (!out_buffer) { j->rc = -3; return nullptr; }
    uint32_t counters[32];
    std::memset(counters, 0, sizeof counters);
    uint64_t since_widen = 0;
    const Codecs& c = codecs();
    for (size_t b = j->first; b < j->count; b += j->step) {
        const Blk& k = j->blk[b];
        if (k.raw < 0 || k.raw > 1024000 + 65536) { j->rc = -4; break; }
        const auto t0 = std::chrono::steady_clock::now();
        if (j->codec == 0) {
            const int got = c.lz4((const char*)k.payload, (char*)out_buffer, k.comp, k.raw);
            if (got <= 0) { j->rc = -5; break; }
        } else {
            const size_t got = c.zstd(out_buffer, (size_t)k.raw, k.payload, (size_t)k.comp);
            if (c.zstd_is_error(got)) { j->rc = -5; break; }
        }
        j->decode_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const uint32_t N = (uint32_t)k.raw >> 1;
        FLAGSTATS_func func = FLAGSTATS_get_function(N);
        (*func)((uint16_t*)out_buffer, N, counters);
        since_widen += N;
        if (since_widen > (1ull << 31)) { /* the reference's uint32_t counters would wrap */
            for (int i = 0; i < 32; ++i) { j->flags[i] += counters[i]; counters[i] = 0; }
            since_widen = 0;
        }
    }
    for (int i = 0; i < 32; ++i) j->flags[i] += counters[i];
    STORM_aligned_free(out_buffer);
    return nullptr;
}

}  // namespace

extern "C" {

/* STORM_pospopcnt_u16 over contiguous ranges on nthreads pthreads; out64 is OVERWRITTEN
 * (zero-then-count like the kernel itself). */
int ref_pospopcnt_mt(const uint16_t* data, uint64_t len, int nthreads, uint64_t* out64, double* seconds)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    PopJob* jobs = (PopJob*)std::calloc((size_t)nthreads, sizeof(PopJob));
    pthread_t* th = (pthread_t*)std::calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { std::free(jobs); std::free(th); return -3; }
    const uint64_t per = (len / (uint64_t)nthreads) & ~511ull;
    uint64_t off = 0;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].base = data + off;
        jobs[t].len = (t == nthreads - 1) ? (len - off) : per;
        off += jobs[t].len;
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        pop_thread(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], nullptr, pop_thread, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int k = 0; k < 16; ++k) out64[k] = 0;
    for (int t = 0; t < nthreads; ++t)
        for (int k = 0; k < 16; ++k) out64[k] += jobs[t].out[k];
    std::free(jobs);
    std::free(th);
    return 0;
}

/* 1 = the codec's runtime library was found (codec 0 = LZ4, 1 = Zstd) */
int ref_codec_available(int codec)
{
    const Codecs& c = codecs();
    return codec == 0 ? (c.lz4 != nullptr) : (c.zstd != nullptr && c.zstd_is_error != nullptr);
}

/* The reference's block loop over a [int32 raw][int32 comp][payload] container held in
 * memory (the file is read before the clock starts, as the page cache would serve it):
 * nthreads = 1 is lz4_decompress() / zstd_decompress() as shipped; nthreads > 1 deals the
 * blocks round-robin to pthreads, each with its own out_buffer and counters[32] (our
 * wrapper, the reference has no threads).  flags64 is ACCUMULATED into; *seconds = wall time
 * of the loop, *decode_seconds = the part of it spent inside the codec (max over threads). */
int ref_container_mt(const void* bytes, uint64_t n_bytes, int codec, int nthreads, uint64_t* flags64,
                     uint64_t* n_records, double* seconds, double* decode_seconds)
{
    if (!ref_codec_available(codec)) return -6;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    const unsigned char* p = (const unsigned char*)bytes;
    size_t n_blk = 0, cap = 1024;
    Blk* blk = (Blk*)std::malloc(cap * sizeof(Blk));
    if (!blk) return -3;
    uint64_t pos = 0, recs = 0;
    while (pos < n_bytes) {
        int32_t hdr[2];
        if (n_bytes - pos < sizeof hdr) { std::free(blk); return -4; }
        std::memcpy(hdr, p + pos, sizeof hdr);
        pos += sizeof hdr;
        if (hdr[0] < 0 || hdr[1] <= 0 || (uint64_t)hdr[1] > n_bytes - pos) { std::free(blk); return -4; }
        if (n_blk == cap) {
            cap *= 2;
            Blk* nb = (Blk*)std::realloc(blk, cap * sizeof(Blk));
            if (!nb) { std::free(blk); return -3; }
            blk = nb;
        }
        blk[n_blk++] = Blk{p + pos, hdr[0], hdr[1]};
        recs += (uint32_t)hdr[0] >> 1;
        pos += (uint64_t)hdr[1];
    }
    ContJob* jobs = (ContJob*)std::calloc((size_t)nthreads, sizeof(ContJob));
    pthread_t* th = (pthread_t*)std::calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !th) { std::free(jobs); std::free(th); std::free(blk); return -3; }
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].blk = blk;
        jobs[t].first = (size_t)t;
        jobs[t].step = (size_t)nthreads;
        jobs[t].count = n_blk;
        jobs[t].codec = codec;
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        cont_thread(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], nullptr, cont_thread, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int rc = 0;
    double dec = 0.0;
    for (int t = 0; t < nthreads; ++t) {
        if (jobs[t].rc) rc = jobs[t].rc;
        if (jobs[t].decode_s > dec) dec = jobs[t].decode_s;
        for (int k = 0; k < 32; ++k) flags64[k] += jobs[t].flags[k];
    }
    if (decode_seconds) *decode_seconds = dec;
    if (n_records) *n_records = recs;
    std::free(jobs);
    std::free(th);
    std::free(blk);
    return rc;
}

}  /* extern "C" */
