/*
 * flagstat_oracle.c -- CPU restatement of the libflagstats flagstat hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libflagstats_b200/,
 * include/) may call, link or import this file.  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * have an independent checker for the CUDA path.
 *
 * Parity is PINNED: oracle/_ref/ holds the unmodified reference compiled from
 * /root/reference (see oracle/Makefile) and tests/test_oracle_vs_reference.py
 * checks every function here against it; tests/golden/ holds the resulting
 * known-answer vectors (KAT A-E) for the GPU box where
 * /root/reference does not exist.
 *
 * All citations are file:line in /root/reference.
 *
 * Differences from the reference that are deliberate:
 *   - 64-bit lengths and 64-bit counters (the reference uses uint32_t for
 *     both, libflagstats.h:170); the *_u32 entry narrows with the same
 *     mod-2^32 wrap the reference's "++" has.
 *   - Written from the rule, not from the SIMD code: one function states the
 *     branchy rule (libflagstats.h:118-142), a second one states the
 *     mask-select formulation (paper/scripts/mask_data.py:29-46,
 *     paper/scripts/expand_data.py:3-10) and the two are cross-checked in
 *     tests.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

/* SAM FLAG bits, numbering as libflagstats.h:69-112 */
enum {
    B_PAIRED = 0, B_PROPER = 1, B_UNMAP = 2, B_MUNMAP = 3, B_REVERSE = 4,
    B_MREVERSE = 5, B_READ1 = 6, B_READ2 = 7, B_SECONDARY = 8, B_QCFAIL = 9,
    B_DUP = 10, B_SUPP = 11, B_PAIR_GOOD = 12, B_SGLTN = 13, B_PAIR_MAP = 14
};

#define BIT(v, b) (((v) >> (b)) & 1u)

/*
 * The normative per-record rule, libflagstats.h:118-142
 * (FLAGSTAT_scalar_update).  f points at 32 counters: [0..15] QC-pass,
 * [16..31] QC-fail.  Bits 12..15 of the input are ignored, exactly as the
 * reference's scalar code ignores them (it only ever tests bits 0..11).
 */
void oracle_flagstat_update(uint16_t v, uint64_t* flags)
{
    uint64_t* f = flags + (BIT(v, B_QCFAIL) ? 16 : 0);           /* :122-123 */
    if (BIT(v, B_QCFAIL)) f[B_QCFAIL] += 1;                      /* :127 */

    if (BIT(v, B_SECONDARY)) {                                   /* :129 */
        f[B_SECONDARY] += 1;
    } else if (BIT(v, B_SUPP)) {                                 /* :130 */
        f[B_SUPP] += 1;
    } else if (BIT(v, B_PAIRED)) {                               /* :131 */
        if (BIT(v, B_PROPER) && !BIT(v, B_UNMAP)) f[B_PAIR_GOOD] += 1; /* :133 */
        if (BIT(v, B_READ1)) f[B_READ1] += 1;                    /* :134 */
        if (BIT(v, B_READ2)) f[B_READ2] += 1;                    /* :135 */
        if (BIT(v, B_MUNMAP) && !BIT(v, B_UNMAP)) f[B_SGLTN] += 1;     /* :136 */
        if (!BIT(v, B_UNMAP) && !BIT(v, B_MUNMAP)) f[B_PAIR_MAP] += 1; /* :137 */
    }
    if (BIT(v, B_UNMAP)) f[B_UNMAP] += 1;                        /* :140 */
    if (BIT(v, B_DUP)) f[B_DUP] += 1;                            /* :141 */
}

/*
 * Scalar convention (libflagstats.h:170-176): accumulate into flags, slot 9
 * (QC-pass record count) is never written.
 */
int oracle_flagstat_scalar_u64(const uint16_t* a, uint64_t n, uint64_t* flags)
{
    for (uint64_t i = 0; i < n; ++i) oracle_flagstat_update(a[i], flags);
    return 0;
}

/*
 * SIMD convention: what FLAGSTAT_sse4 / _avx2 / _avx512 -- the kernels
 * FLAGSTATS_u16 dispatches to for n >= 256 -- leave in CORE19 + slot 9:
 * the scalar counters plus  flags[9] += len - (flags[25] - start_qc)
 * (libflagstats.h:185,429 / :968,1212 / :1647,1843).  This is the contract
 * FLAGSTAT_cuda follows ("Slot 9").
 */
int oracle_flagstat_simd_u64(const uint16_t* a, uint64_t n, uint64_t* flags)
{
    const uint64_t start_qc = flags[16 + B_QCFAIL];
    oracle_flagstat_scalar_u64(a, n, flags);
    flags[B_QCFAIL] += n - (flags[16 + B_QCFAIL] - start_qc);
    return 0;
}

/* The reference's own ABI: uint32_t len, uint32_t counters (wraps mod 2^32). */
int oracle_flagstat_simd_u32(const uint16_t* a, uint32_t n, uint32_t* flags)
{
    uint64_t t[32];
    memset(t, 0, sizeof t);
    oracle_flagstat_simd_u64(a, n, t);
    for (int i = 0; i < 32; ++i) flags[i] += (uint32_t)t[i];
    return 0;
}

int oracle_flagstat_scalar_u32(const uint16_t* a, uint32_t n, uint32_t* flags)
{
    uint64_t t[32];
    memset(t, 0, sizeof t);
    oracle_flagstat_scalar_u64(a, n, t);
    for (int i = 0; i < 32; ++i) flags[i] += (uint32_t)t[i];
    return 0;
}

/*
 * Mask-select formulation.  Returns the 16-bit word whose positional popcount
 * (split by the 0x200 bit) equals the rule above:
 *   bit12 = PAIRED & PROPER & ~UNMAP, bit13 = PAIRED & MUNMAP & ~UNMAP,
 *   bit14 = PAIRED & ~UNMAP & ~MUNMAP      (paper/scripts/expand_data.py:3-10)
 *   keep {2,9,10} always; +8 if SECONDARY; else +11 if SUPPLEMENTARY;
 *   else +{6,7,12,13,14} if PAIRED        (paper/scripts/mask_data.py:29-46)
 */
uint16_t oracle_mask_select(uint16_t v)
{
    uint32_t x = v & 0x0FFFu;
    const uint32_t p = BIT(x, B_PAIRED), u = BIT(x, B_UNMAP), m = BIT(x, B_MUNMAP);
    x |= (p & BIT(x, B_PROPER) & (u ^ 1u)) << B_PAIR_GOOD;
    x |= (p & m & (u ^ 1u)) << B_SGLTN;
    x |= (p & (u ^ 1u) & (m ^ 1u)) << B_PAIR_MAP;

    uint32_t keep = (1u << B_UNMAP) | (1u << B_QCFAIL) | (1u << B_DUP);
    if (BIT(x, B_SECONDARY)) keep |= 1u << B_SECONDARY;
    else if (BIT(x, B_SUPP)) keep |= 1u << B_SUPP;
    else if (p) keep |= (1u << B_READ1) | (1u << B_READ2) | (1u << B_PAIR_GOOD) |
                        (1u << B_SGLTN) | (1u << B_PAIR_MAP);
    return (uint16_t)(x & keep);
}

/* flagstat through mask-select + positional popcount; SIMD convention. */
int oracle_flagstat_maskselect_u64(const uint16_t* a, uint64_t n, uint64_t* flags)
{
    uint64_t n_fail = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint16_t y = oracle_mask_select(a[i]);
        const int fail = BIT(y, B_QCFAIL);
        uint64_t* f = flags + (fail ? 16 : 0);
        n_fail += (uint64_t)fail;
        for (int j = 0; j < 15; ++j)
            if (j != B_QCFAIL) f[j] += BIT(y, j);
    }
    flags[16 + B_QCFAIL] += n_fail;
    flags[B_QCFAIL] += n - n_fail;
    return 0;
}

/*
 * The reference benchmark's samtools-style accumulator: bam_flagstat_t
 * (benchmark/flagstats.cpp:43-49) filled by flagstat_loop (:51-71), the caller
 * that prints the report (:577-588).  s[] is that struct seen as 13 pairs of
 * long long, [pass, fail] each, in declaration order:
 *   0 n_reads  1 n_mapped  2 n_pair_all  3 n_pair_map  4 n_pair_good  5 n_sgltn
 *   6 n_read1  7 n_read2   8 n_dup  9 n_diffchr  10 n_diffhigh  11 n_secondary
 *   12 n_supp
 * n_diffchr / n_diffhigh need RNAME / MAPQ and are never written (:577-588 has
 * them commented out).  Accumulates.
 */
enum { S_READS = 0, S_MAPPED, S_PAIR_ALL, S_PAIR_MAP, S_PAIR_GOOD, S_SGLTN, S_READ1, S_READ2,
       S_DUP, S_DIFFCHR, S_DIFFHIGH, S_SECONDARY, S_SUPP, S_FIELDS };

int oracle_samtools_loop(const uint16_t* a, uint64_t n, long long* s)
{
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t c = a[i];
        const int w = BIT(c, B_QCFAIL) ? 1 : 0;                               /* :52 */
        s[2 * S_READS + w] += 1;                                             /* :53 */
        if (BIT(c, B_SECONDARY)) {                                           /* :54 */
            s[2 * S_SECONDARY + w] += 1;
        } else if (BIT(c, B_SUPP)) {                                         /* :56 */
            s[2 * S_SUPP + w] += 1;
        } else if (BIT(c, B_PAIRED)) {                                       /* :58 */
            s[2 * S_PAIR_ALL + w] += 1;                                      /* :59 */
            if (BIT(c, B_PROPER) && !BIT(c, B_UNMAP)) s[2 * S_PAIR_GOOD + w] += 1;   /* :60 */
            if (BIT(c, B_READ1)) s[2 * S_READ1 + w] += 1;                    /* :61 */
            if (BIT(c, B_READ2)) s[2 * S_READ2 + w] += 1;                    /* :62 */
            if (BIT(c, B_MUNMAP) && !BIT(c, B_UNMAP)) s[2 * S_SGLTN + w] += 1;       /* :63 */
            if (!BIT(c, B_UNMAP) && !BIT(c, B_MUNMAP)) s[2 * S_PAIR_MAP + w] += 1;   /* :64-66 */
        }
        if (!BIT(c, B_UNMAP)) s[2 * S_MAPPED + w] += 1;                      /* :68 */
        if (BIT(c, B_DUP)) s[2 * S_DUP + w] += 1;                            /* :69 */
    }
    return 0;
}

/*
 * Raw positional popcount: zero out[0..15], then out[j] = #records with bit j
 * set (libalgebra.h:3497-3498 memset, :565-574 naive kernel).  64-bit variant
 * for lengths whose counts exceed 2^32.
 */
int oracle_pospopcnt_u16_u64(const uint16_t* a, uint64_t n, uint64_t* out)
{
    memset(out, 0, 16 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i)
        for (int j = 0; j < 16; ++j) out[j] += BIT(a[i], j);
    return 0;
}

int oracle_pospopcnt_u16(const uint16_t* a, size_t n, uint32_t* out)
{
    uint64_t t[16];
    oracle_pospopcnt_u16_u64(a, n, t);
    for (int j = 0; j < 16; ++j) out[j] = (uint32_t)t[j];
    return 0;
}

/* ------------------------------------------------------------------------
 * Synthetic FLAG columns as pure functions of the GLOBAL record index, so a
 * shard, the GPU and this oracle regenerate identical data without moving it
 * (SURVEY.md section 8d).  The device twins live in
 * libflagstats_b200/csrc/synth.cuh; tests compare the two byte for byte.
 * ---------------------------------------------------------------------- */

static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Four records per hash: record i takes 16 bits of hash(seed, i/4). */
static inline uint16_t synth_uniform_at(uint64_t seed, uint64_t i, uint16_t mask)
{
    const uint64_t h = mix64(seed + ((i >> 2) + 1) * 0x9E3779B97F4A7C15ull);
    return (uint16_t)((h >> (16 * (i & 3))) & mask);
}

/* U(0, mask) -- the distribution of benchmark/generate.cpp:11 and
 * benchmark/inmemory.cpp:113 when mask = 0x0FFF. */
void oracle_synth_uniform(uint16_t* out, uint64_t start, uint64_t n,
                          uint64_t seed, uint16_t mask)
{
    for (uint64_t k = 0; k < n; ++k) out[k] = synth_uniform_at(seed, start + k, mask);
}

/*
 * HiSeqX-shaped column: KAT-E of SURVEY.md section 8c.  Exactly reproduces the
 * FLAG-derivable lines of the samtools output quoted in README.md:179-191 when
 * n == 824,541,892 and start == 0.
 */
#define HISEQX_N 824541892ull
#define HISEQX_M 509594915ull /* ~0.618*N, odd, coprime with N = 2^2*37*5571229 (checked in tests) */
#define HISEQX_NCAT 16

static const uint16_t hiseqx_flag[HISEQX_NCAT] = {
    99, 147, 83, 163,       /* proper pairs */
    97, 145,                /* both mapped, not proper */
    73, 137,                /* singletons (mate unmapped) */
    133, 69,                /* unmapped, mate mapped */
    77, 141,                /* both unmapped */
    2113, 2177, 2129, 2193  /* supplementary */
};
static const uint64_t hiseqx_count[HISEQX_NCAT] = {
    195271471ull, 195271471ull, 195271471ull, 195271471ull,
    8432503ull, 8432503ull,
    1019443ull, 1019442ull,
    1019443ull, 1019442ull,
    8559802ull, 8559802ull,
    1348407ull, 1348407ull, 1348407ull, 1348407ull
};

static inline uint16_t synth_hiseqx_at(uint64_t i, uint64_t seed, uint32_t qcfail_ppm)
{
    const uint64_t j = ((i % HISEQX_N) * HISEQX_M) % HISEQX_N;
    uint64_t acc = 0;
    uint16_t v = 0;
    for (int c = 0; c < HISEQX_NCAT; ++c) {
        acc += hiseqx_count[c];
        if (j < acc) { v = hiseqx_flag[c]; break; }
    }
    if (qcfail_ppm) {
        const uint64_t h = mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull);
        if ((uint32_t)(h % 1000000ull) < qcfail_ppm) v |= 0x200;
    }
    return v;
}

void oracle_synth_hiseqx(uint16_t* out, uint64_t start, uint64_t n,
                         uint64_t seed, uint32_t qcfail_ppm)
{
    for (uint64_t k = 0; k < n; ++k) out[k] = synth_hiseqx_at(start + k, seed, qcfail_ppm);
}

uint64_t oracle_hiseqx_n(void) { return HISEQX_N; }
uint64_t oracle_hiseqx_m(void) { return HISEQX_M; }
