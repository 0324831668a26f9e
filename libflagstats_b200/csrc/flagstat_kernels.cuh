// flagstat_kernels.cuh -- the sm_100a flagstat / pospopcnt kernels.
//
// What is computed is the reference's per-record rule
// (libflagstats.h:118-142, FLAGSTAT_scalar_update) plus the QC-pass record
// count convention of the kernels FLAGSTATS_u16 dispatches to (:429,1212,1843).
// How it is computed is B200-specific and shares nothing with the reference's
// SSE/AVX code:
//
//   * two FLAG records stay packed in one 32-bit register from the 128-bit load
//     to the counters -- nothing is ever unpacked;
//   * the samtools if / else-if chain is folded into ONE word Y per register by
//     "mask select": two fp16x2 equality compares on the masked word produce
//     full-halfword masks directly (they run on the half/FMA pipe and leave the
//     integer ALU pipe, the real bottleneck, to the LOP3s), the keep-mask is one
//     LOP3 and applying it a second one;
//   * Y is fed to a bit-sliced carry-save counter (bitcounter.cuh); a second
//     counter takes Y & failmask.  QC-pass = all - fail, so the pass side is
//     never materialised, and the second counter is skipped for every warp
//     batch that holds no QC-fail record (always, in real data);
//   * position -> counter mapping happens once per CTA at the very end.
//
// Y layout per 16-bit record (positions not listed carry garbage that is
// simply never read back):
//   0  G  = PAIRED & ~SEC & ~SUPP & ~UNMAP        (n_pair_map = N(G) - N(pos 3))
//   1  PROPER & G                                  -> slot 12 n_pair_good
//   2  UNMAP                                       -> slot 2
//   3  MUNMAP & G                                  -> slot 13 n_sgltn
//   6  READ1 & K,  7  READ2 & K   (K = PAIRED & ~SEC & ~SUPP) -> slots 6, 7
//   8  SECONDARY                                   -> slot 8
//   9  QCFAIL                                      -> slot 25 (and 9 = n - fail)
//   10 DUP                                         -> slot 10
//   11 SUPPLEMENTARY & ~SECONDARY                  -> slot 11
// Bits 12..15 of the input are ignored, as the scalar reference ignores them.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "bitcounter.cuh"

namespace fsb200 {

constexpr int kThreads = 256;              // threads per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kU = 4;                      // 128-bit loads per thread per batch
constexpr int kVecPerBatch = kThreads * kU;  // uint4 per CTA batch (16 KiB)

// kSamtools = kFlagstat plus the exact "paired in sequencing" count (n_pair_all of the
// reference benchmark's flagstat_loop, benchmark/flagstats.cpp:58-59) in slots 0 / 16.
enum Mode { kFlagstat = 0, kPospopcnt = 1, kSamtools = 2 };

// ---------------------------------------------------------------------------
// mask select
// ---------------------------------------------------------------------------

// per-halfword equality of two packed fp16 pairs -> 0xFFFF / 0x0000 per half.
// No .ftz: the operands are subnormal bit patterns and must compare exactly.
__device__ __forceinline__ uint32_t eq2_mask(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t ne2_mask(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// explicit LOP3 with an immediate truth table (a = 0xF0, b = 0xCC, c = 0xAA);
// spelled out so ptxas keeps the intended 3-input grouping
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// Variant H (default): half-pipe compares.  Four ALU-pipe LOP3 per packed pair;
// the two compares and the shift run on the half / FMA pipes.
__device__ __forceinline__ uint32_t mask_select_h(uint32_t w)
{
    // PAIRED(0) UNMAP(2) SEC(8) SUPP(11) of both records
    const uint32_t q = w & 0x09050905u;
    const uint32_t gm = eq2_mask(q, 0x00010001u);  // G: paired, mapped, primary
    const uint32_t xm = eq2_mask(q, 0x00050005u);  // K & UNMAP: paired, unmapped, primary
    // G keeps everything, K & UNMAP keeps READ1/READ2 only:  e = gm | (xm & 0x00C0)
    const uint32_t e = lop3<0xF8>(gm, xm, 0x00C000C0u);
    // SUPP only counts when not SEC:  wf = w & ~((w << 3) & 0x0800)
    const uint32_t wf = lop3<0x70>(w, w << 3, 0x08000800u);
    // always kept: UNMAP SEC QCFAIL DUP SUPP':  y = wf & (e | 0x0F04)
    return lop3<0xE0>(wf, e, 0x0F040F04u);
}

// QCFAIL (bit 9) of each record spread over its whole halfword: the shift (FMA
// pipe) puts the bit in the sign position of the odd bytes, PRMT's
// sign-replicate mode (selector nibbles 0x9 / 0xB) does the broadcast.
// Same, for a warp batch in which no record has SECONDARY set (all of real-world
// HiSeqX data, README.md:180): SUPP needs no fix-up, one LOP3 and the shift less.
__device__ __forceinline__ uint32_t mask_select_h_nosec(uint32_t w)
{
    const uint32_t q = w & 0x09050905u;
    const uint32_t gm = eq2_mask(q, 0x00010001u);
    const uint32_t xm = eq2_mask(q, 0x00050005u);
    const uint32_t e = lop3<0xF8>(gm, xm, 0x00C000C0u);
    return lop3<0xE0>(w, e, 0x0F040F04u);
}

// Variant T: the two class tests as fp16x2 "tent" functions on the half/FMA pipe
// instead of two HSET2 compares on the integer ALU pipe.  q = w & 0x0905 per
// half is a small non-negative fp16 value (a subnormal when SUPPLEMENTARY is
// clear); with ulp = 2^-24 (bit pattern 0x0001)
//     u = sat(ulp - |q - 1 ulp|)  is 1 ulp iff q == 0x0001 (G),      else +0
//     v = sat(ulp - |q - 5 ulp|)  is 1 ulp iff q == 0x0005 (K&UNMAP), else +0
// (HADD2.SAT clamps below at +0; |.| and - are free operand modifiers), and
//     e = 203 * u + 192 * v   has the bit pattern 0x00CB, 0x00C0 or 0x0000:
// exactly the bits G / K&UNMAP records keep in their low byte.  All of this is
// exact: the operands are integers < 1024 in units of ulp, or so large that the
// tents are 0.  Six half-pipe instructions replace three ALU-pipe ones.
__device__ __forceinline__ uint32_t class_bits_t(uint32_t w)
{
    const uint32_t q = w & 0x09050905u;
    uint32_t t1, t5, u, v, m, e;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(t1) : "r"(q), "r"(0x80018001u));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(t5) : "r"(q), "r"(0x80058005u));
    asm("{ .reg .b32 a; abs.f16x2 a, %1; neg.f16x2 a, a; add.rn.sat.f16x2 %0, a, %2; }"
        : "=r"(u) : "r"(t1), "r"(0x00010001u));
    asm("{ .reg .b32 a; abs.f16x2 a, %1; neg.f16x2 a, a; add.rn.sat.f16x2 %0, a, %2; }"
        : "=r"(v) : "r"(t5), "r"(0x00010001u));
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(m) : "r"(v), "r"(0x5A005A00u));             // 192.0
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(e) : "r"(u), "r"(0x5A585A58u), "r"(m));  // 203.0
    return e;
}

__device__ __forceinline__ uint32_t mask_select_t(uint32_t w)
{
    const uint32_t e = class_bits_t(w);
    const uint32_t wf = lop3<0x70>(w, w << 3, 0x08000800u);
    return lop3<0xE0>(wf, e, 0x0F040F04u);
}

__device__ __forceinline__ uint32_t mask_select_t_nosec(uint32_t w)
{
    return lop3<0xE0>(w, class_bits_t(w), 0x0F040F04u);
}

__device__ __forceinline__ uint32_t fail_mask_h(uint32_t w)
{
    // inline PTX: __byte_perm() documents selector bit 3 as ignored, prmt.b32 does not
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w << 6), "r"(0u), "r"(0xBB99u));
    return d;
}

// Variant F (default since round 2): the keep-mask is BUILT on the FMA pipe.
// The two class tests come out of HSET2 as fp16 1.0 / 0.0 (.BF form) and HFMA2
// add the bits each class keeps to the always-kept base, in units of the base's
// own ulp (2^-22; 0x0F04 = 1796 ulp, exponent field 3):
//     k = 0x0F04 + [K & UNMAP] * 0xC0 + [G] * 0xCB   ->  0x0FC4 / 0x0FCF / 0x0F04
// (0xC0 ulp = fp16 subnormal 0x0300, 0xCB ulp = 0x032C; every sum is < 2048 ulp,
// so the additions are exact and the bit patterns are the masks themselves).
// SUPPLEMENTARY must not count for SECONDARY records (the reference's else-if):
// q = w & 0x0905 read as fp16 is >= 0x0900 exactly when both bits are set, the
// next smaller possible value is 0x0805, and
//     z = sat(q * 49152 - 6.03125)   is 1.0 for the former and 0 for everything else
// (0x0900 -> 7.5 - 6.03, 0x0805 -> 6.0293 - 6.03125 < 0).  Such a record is never
// in class G or K, so its k is the base, and k - z * 1347 ulp = 449 ulp = 0x0704:
// the base with bit 11 cleared.  One AND then applies everything:  y = w & k.
// ALU pipe: AND, HSET2, HSET2, AND = 4 (was 5, or 6 with a SECONDARY record in the
// batch); FMA pipe: 4 HFMA2.  y has bits 4, 5, 12..15 CLEAR (the HSET2-mask form
// let input bits through for G records).  The pipe microbenchmark shows HFMA2 /
// HMUL2 issue for free beside LOP3 up to 1 per 2 ALU instructions
// (profiles/r2b_pipe_microbench.txt).
__device__ __forceinline__ uint32_t mask_select_f(uint32_t w)
{
    const uint32_t q = w & 0x09050905u;
    uint32_t g1, x1, z, k;
    asm("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(g1) : "r"(q), "r"(0x00010001u));
    asm("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(x1) : "r"(q), "r"(0x00050005u));
    asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(z) : "r"(q), "r"(0x7A007A00u), "r"(0xC608C608u));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(x1), "r"(0x03000300u), "r"(0x0F040F04u));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(g1), "r"(0x032C032Cu), "r"(k));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(z), "r"(0x8D438D43u), "r"(k));
    return w & k;
}

// kSamtools form of the above: position 4 becomes the indicator of class K (paired,
// primary) = one count of n_pair_all per record (benchmark/flagstats.cpp:58-59).  The
// class constants gain bit 4 (0xD0 / 0xDB ulp: 0x0FD4 / 0x0FDF, still < 2048 ulp, so
// still exact) and the final AND sees the input with bit 4 forced on:
//     y = k & (w | 0x0010)        one LOP3 (0xE0), like the plain AND it replaces
// so K records carry a 1 at position 4 whatever their REVERSE bit says and all other
// records (k = 0x0F04 / 0x0704) a 0.  Same instruction count as mask_select_f.
template <bool HAS_SEC>
__device__ __forceinline__ uint32_t mask_select_fx(uint32_t w)
{
    const uint32_t q = w & 0x09050905u;
    uint32_t g1, x1, k;
    asm("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(g1) : "r"(q), "r"(0x00010001u));
    asm("set.eq.f16x2.f16x2 %0, %1, %2;" : "=r"(x1) : "r"(q), "r"(0x00050005u));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(x1), "r"(0x03400340u), "r"(0x0F040F04u));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(g1), "r"(0x036C036Cu), "r"(k));
    if (HAS_SEC) {  // SUPPLEMENTARY does not count for SECONDARY records (see mask_select_f)
        uint32_t z;
        asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(z) : "r"(q), "r"(0x7A007A00u), "r"(0xC608C608u));
        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(k) : "r"(z), "r"(0x8D438D43u), "r"(k));
    }
    return lop3<0xE0>(k, w, 0x00100010u);  // k & (w | 0x0010)
}

// QC-fail gating of a variant-F word without an integer mask: bit 9 of the record,
// read as fp16, is the subnormal 2^-15; times 2^15 it is 1.0 or 0.0, and y * 1.0
// is y exactly (y <= 0x0FCF is a finite non-negative fp16 -- this needs the CLEAN
// bits 12..15 of mask_select_f), y * 0.0 = +0.  One ALU instruction (the AND)
// and two HMUL2 instead of shift + PRMT + AND.
__device__ __forceinline__ uint32_t fail_gate_f(uint32_t w, uint32_t y)
{
    const uint32_t f = w & 0x02000200u;
    uint32_t f1, yf;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(f1) : "r"(f), "r"(0x78007800u));
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(yf) : "r"(y), "r"(f1));
    return yf;
}

// Variant I: integer-only formulation (A/B reference for the one above and a
// safety net should a future toolchain change fp16 subnormal semantics).
__device__ __forceinline__ uint32_t mask_select_i(uint32_t w)
{
    const uint32_t k0 = w & ~(w >> 8) & ~(w >> 11);      // bit0 of each half: K
    const uint32_t g0 = k0 & ~(w >> 2) & 0x00010001u;    // bit0: G
    const uint32_t e = (k0 & 0x00010001u) * 0x00C0u + g0 * 0x000Bu;  // K->{6,7} G->{0,1,3}
    const uint32_t s11 = (w << 3) & 0x08000800u;
    const uint32_t wf = w & ~s11;
    return wf & (e | 0x0F040F04u);
}

__device__ __forceinline__ uint32_t fail_mask_i(uint32_t w)
{
    return ((w >> 9) & 0x00010001u) * 0xFFFFu;
}

template <int VARIANT>
__device__ __forceinline__ uint32_t mask_select(uint32_t w)
{
    return VARIANT == 1 ? mask_select_i(w) : mask_select_h(w);
}

template <int VARIANT>
__device__ __forceinline__ uint32_t fail_mask(uint32_t w)
{
    return VARIANT == 1 ? fail_mask_i(w) : fail_mask_h(w);
}

// ---------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------

// streaming 128-bit load: read-only path, no L1 allocation (every byte is used
// exactly once)
__device__ __forceinline__ uint4 ld_stream(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------
// per-thread state and the batch step
// ---------------------------------------------------------------------------

using Counter = BitCounter<4, 6>;  // 14 planes: 16 * (2^10 - 2) words per epoch

template <int MODE, int VARIANT>
struct Lanes {
    Counter all;
    Counter fail;   // unused in pospopcnt mode (dead-code eliminated)
    uint32_t nfail;  // batches absorbed by `fail` this epoch (warp-uniform)

    __device__ __forceinline__ void clear()
    {
        all.clear();
        if (MODE == kFlagstat) fail.clear();
        nfail = 0u;
    }

    // one batch = 16 packed words; b = batches absorbed by `all` this epoch
    __device__ __forceinline__ void step(const uint32_t (&w)[16], uint32_t b)
    {
        if (MODE == kPospopcnt) {
            all.absorb16(w, b);
            return;
        }
        // Warp-uniform view of the batch: OR of all 512 packed words (REDUX.OR).
        uint32_t any = w[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) any |= w[i];
        const uint32_t wany = __reduce_or_sync(0xffffffffu, any);

        uint32_t y[16];
        if (VARIANT == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = mask_select_i(w[i]);
        } else if ((wany & 0x01000100u) == 0u) {  // no SECONDARY record in this warp batch
#pragma unroll
            for (int i = 0; i < 16; ++i)
                y[i] = VARIANT == 2 ? mask_select_t_nosec(w[i]) : mask_select_h_nosec(w[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = VARIANT == 2 ? mask_select_t(w[i]) : mask_select_h(w[i]);
        }
        all.absorb16(y, b);
        // Any QCFAIL record in this warp batch?  In real data there is none and
        // the whole second counter is skipped.
        if ((wany & 0x02000200u) != 0u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] &= fail_mask<VARIANT>(w[i]);
            fail.absorb16(y, nfail);
            ++nfail;
        }
    }
};

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------

__device__ __forceinline__ void unpack4(const uint4 (&v)[kU], uint32_t (&w)[16])
{
#pragma unroll
    for (int u = 0; u < kU; ++u) {
        w[4 * u + 0] = v[u].x;
        w[4 * u + 1] = v[u].y;
        w[4 * u + 2] = v[u].z;
        w[4 * u + 3] = v[u].w;
    }
}

__device__ __forceinline__ void load_batch(uint4 (&v)[kU], const uint4* p)
{
#pragma unroll
    for (int u = 0; u < kU; ++u) v[u] = ld_stream(p + u * kThreads);
}

// position (within a 16-bit record) -> primary output slot, -1 = not a counter
template <int MODE>
__device__ __forceinline__ int slot_of_position(int p)
{
    if (MODE == kSamtools && p == 4) return 0;  // K -> n_pair_all
    switch (p) {
        case 0: return 14;  // G; position 3 is subtracted below
        case 1: return 12;
        case 2: return 2;
        case 3: return 13;
        case 6: return 6;
        case 7: return 7;
        case 8: return 8;
        case 10: return 10;
        case 11: return 11;
        default: return -1;
    }
}

// Lanes 0..15 of one warp hold the CTA totals of tree position `lane`
// (a = all records, f = QC-fail records).  Map positions to the reference's
// counter slots and add them to out[] with 64-bit atomics; differences are
// taken mod 2^64, which is exact for sums of non-negative counts.
template <int MODE>
__device__ __forceinline__ void emit_counters(unsigned long long* __restrict__ out, uint32_t lane,
                                              unsigned long long a, unsigned long long f,
                                              uint64_t n)
{
    if (lane >= 16) return;
    if (MODE == kPospopcnt) {
        if (a) atomicAdd(out + lane, a);
        return;
    }
    const int slot = slot_of_position<MODE>((int)lane);
    if (slot >= 0) {
        if (a - f) atomicAdd(out + slot, a - f);
        if (f) atomicAdd(out + 16 + slot, f);
        if (lane == 3) {  // n_pair_map = N(G) - N(G & MUNMAP)
            if (a - f) atomicAdd(out + 14, 0ull - (a - f));
            if (f) atomicAdd(out + 30, 0ull - f);
        }
    } else if (lane == 9) {  // a == number of QC-fail records
        unsigned long long pass = 0ull - a;
        if (blockIdx.x == 0) pass += n;  // slot 9 = n - n_fail, libflagstats.h:429
        if (a) atomicAdd(out + 25, a);
        if (pass) atomicAdd(out + 9, pass);
    }
}

// ---------------------------------------------------------------------------
// fused counter exchange (multi-GPU): count -> push -> wait -> sum in ONE kernel
// ---------------------------------------------------------------------------
//
// Every rank (GPU) owns one exchange buffer in its own HBM; all ranks map all
// buffers (CUDA IPC across processes, peer access inside one process).  When
// xa.world != 0 the CTAs of a launch add their per-CTA totals into xa.acc
// instead of the caller's counters; the CTA that draws the last ticket then
//   1. takes the rank's 32 totals out of xa.acc (and re-zeroes it),
//   2. stores them into slot [epoch & 1][rank] of EVERY rank's buffer with plain
//      stores over NVLink, each 64-bit counter as TWO self-validating 8-byte words
//      {32 bits of the counter, 32-bit tag of the epoch},
//   3. reads the `world` slots of its OWN buffer until every word carries this epoch's
//      tag, adds them and writes the global counters.
// A word is valid exactly when its tag matches, and an aligned 8-byte store is never
// torn, so there is no flag, and no fence between data and flag: the pushing CTA fires
// its stores and is done (round 1 pushed 32 plain words, __threadfence_system(), then a
// flag per peer -- the fence waits for the acknowledgement of every remote store, and with
// overlapped steps that wait sat on the critical path: the next kernel's CTA that inherits
// the exchanging CTA's SM slot starts late by the length of the exchange, draws the last
// ticket itself, and the delay adds up once per step; 7 us at 8 GPUs).
// The exchange is 512 bytes per peer; there is no separate collective launch.
//
// Deferred collection (FLAGSTAT_cuda_device_allreduce_deferred) splits step 3 off: a launch
// only does 1-2 for its own epoch, and FIRST finishes step 3 of the epoch its predecessor on
// the handle left pending (or FLAGSTAT_cuda_xchg_collect does, as a one-warp kernel).  A rank
// then never waits for the peers' CURRENT step, only for their previous one: it may run one
// whole step ahead of the slowest rank, so per-step jitter between GPUs no longer adds up
// (with step 3 in the same launch every step ends with the slowest rank of THAT step).  The
// parity double buffer is still sufficient: rank r overwrites slot [e & 1] with epoch e + 2
// only after it has collected e + 1, and a peer pushes e + 1 only after it has
// collected (i.e. finished reading) e.  tests/test_exchange_protocol_model.py runs both
// orders, and mixtures of them, under every interleaving.
//
// Overlapped steps (opt-in, FLAGSTAT_cuda_xchg_set_overlap): the launch carries the
// programmatic-stream-serialization attribute, every CTA executes
// griddepcontrol.launch_dependents at its start and griddepcontrol.wait only at the top
// of its epilogue.  The CTAs of step k+1 then take the SM slots the CTAs of step k free
// and stream their input while the last CTA of k is still waiting for its peers: the
// exchange latency, the kernel tail and the slowest rank's skew of one step hide behind
// the counting of the next.  Everything a step PUBLISHES (accumulator, ticket, peer
// slots, out[]) happens after the wait, i.e. after the previous kernel of the stream has
// completed and flushed, so the protocol above is unchanged; only the input column is
// read early, which is why the caller has to opt in (it must not be produced by the
// kernel launched just before on the same stream).  Slots are double-buffered
// by epoch parity: a rank can start epoch e+2 only after every rank finished
// epoch e (it needs their e+1 data to finish e+1), so a slot is never rewritten
// while a peer may still read it.
constexpr int kMaxRanks = 16;
constexpr int kXchgSlotWords = 64;  // 32 counters x {low half | tag, high half | tag}
// layout of one exchange buffer, in 8-byte words
constexpr int kXchgSlots = 0;                                           // [2][kMaxRanks][64]
constexpr int kXchgAcc = kXchgSlots + 2 * kMaxRanks * kXchgSlotWords;    // [32] per-launch accumulator
constexpr int kXchgTicket = kXchgAcc + 32;                               // [1]
constexpr int kXchgErr = kXchgTicket + 1;                                // [1] != 0 after a timeout
constexpr int kXchgWords = kXchgErr + 1;

struct XchgArgs {
    unsigned long long* buf[kMaxRanks];  // buf[r] = rank r's exchange buffer as mapped HERE
    unsigned long long epoch;            // 1, 2, 3, ... (same on every rank for one collective call)
    unsigned long long timeout_ns;       // give up waiting for peers after this long
    int rank;
    int world;                           // 0 = no exchange: plain accumulate into out
    int accumulate;                      // 1: out[i] += total, 0: out[i] = total
    // Deferred collection (FLAGSTAT_cuda_device_allreduce_deferred): this launch only PUSHES its
    // totals; the launch (or FLAGSTAT_cuda_xchg_collect) that follows waits for the peers' totals
    // of this epoch and writes the global counters.  A launch therefore first collects the epoch
    // its predecessor left pending, then pushes its own.
    int deferred;                        // 1: push only
    int prev_accumulate;
    int prev_nout;                       // 16 / 32 counters pending
    unsigned long long prev_epoch;       // != 0: collect this epoch into prev_out first
    unsigned long long* prev_out;
    unsigned long long* host_err;        // mapped host word, receives the epoch of a timeout (may be null)
    // launch geometry worked out on the host (64-bit divisions cost a CTA ~0.3 us at its start):
    // NB full CTA batches = nb_q * gridDim.x + nb_r
    unsigned long long nb_q;
    unsigned int nb_r;
    int pdl;                             // launched with programmatic stream serialization
    unsigned long long* dyn;             // flagstat_kernel_dyn: this launch's {counter, done tickets, owner} slot
    unsigned int dyn_tag;                //   ... and the tag that marks the slot as this launch's
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Programmatic dependent launch; both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior_grids()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Tag of an epoch inside the packed words: never 0 (a zeroed buffer holds no valid word), and
// different for e and e + 2, the two epochs that share a parity slot.
__device__ __forceinline__ unsigned long long xchg_tag(unsigned long long e)
{
    return (e % 0xFFFFFFFFull + 1ull) << 32;
}
__device__ __forceinline__ void st_relaxed_sys_v2(unsigned long long* p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// acquire: whatever this thread stores afterwards (its own push of the next epoch) is ordered behind the read
__device__ __forceinline__ void ld_acquire_sys_v2(const unsigned long long* p, unsigned long long& a,
                                                  unsigned long long& b)
{
    asm volatile("ld.acquire.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Read (all lanes of one warp, lane = counter) every rank's totals of epoch `e` out of this rank's
// buffer, waiting for words that have not landed yet, add them and write `nout` global counters
// to dst.  false = a peer never delivered (or the exchange had already failed): dst untouched,
// error recorded here, on the host and on every peer, so that nobody keeps waiting for a rank
// that has given up.
__device__ __forceinline__ bool xchg_collect(unsigned long long* __restrict__ dst, uint32_t lane,
                                             const XchgArgs& xa, unsigned long long e, int accumulate,
                                             uint32_t nout, unsigned long long failed_before = 0ull)
{
    unsigned long long* mine = xa.buf[xa.rank];
    const int world = xa.world;
    const uint32_t par = (uint32_t)(e & 1ull);
    const unsigned long long tag = xchg_tag(e);
    constexpr unsigned long long kTagMask = 0xFFFFFFFF00000000ull;
    const unsigned long long* w = mine + kXchgSlots + par * kMaxRanks * kXchgSlotWords + 2u * (lane & 31u);
    bool ok = true;
    unsigned long long v = 0ull;
    if (lane < nout) {
        unsigned long long old = 0ull;
        if (accumulate) old = dst[lane];
        // four ranks at a time: eight words in flight per lane, one memory round trip when
        // the peers' words are already there (they normally are: deferred collection)
        for (int r0 = 0; r0 < world && ok; r0 += 4) {
            unsigned long long lo[4], hi[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                lo[k] = hi[k] = tag;
                if (r0 + k < world) ld_acquire_sys_v2(w + (r0 + k) * kXchgSlotWords, lo[k], hi[k]);
            }
            if (failed_before) {  // (looked at only now: its load travels together with the ones above)
                ok = false;
                break;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (r0 + k >= world) continue;
                if (((lo[k] ^ tag) | (hi[k] ^ tag)) & kTagMask) {  // not there yet: poll this one
                    const unsigned long long t0 = global_timer_ns();
                    uint32_t spins = 0;
                    do {
                        if ((++spins & 63u) == 0u &&
                            (global_timer_ns() - t0 > xa.timeout_ns || ld_relaxed_sys(mine + kXchgErr) != 0ull)) {
                            ok = false;
                            break;
                        }
                        ld_acquire_sys_v2(w + (r0 + k) * kXchgSlotWords, lo[k], hi[k]);
                    } while (((lo[k] ^ tag) | (hi[k] ^ tag)) & kTagMask);
                }
                v += (lo[k] & 0xFFFFFFFFull) | (hi[k] << 32);
            }
        }
        v += old;
    }
    if (!__all_sync(0xffffffffu, ok)) {
        const unsigned long long code = failed_before ? failed_before : e;  // the first failure keeps its epoch
        if ((int)lane < world) st_relaxed_sys(xa.buf[lane] + kXchgErr, code);  // own buffer included
        if (lane == 0 && xa.host_err) st_relaxed_sys(xa.host_err, code);
        __threadfence_system();
        return false;
    }
    if (lane < nout) dst[lane] = v;
    return true;
}

// Executed by warp 0 of the CTA that drew the last ticket of a launch.
template <int MODE>
__device__ __noinline__ void xchg_last_cta(unsigned long long* __restrict__ out, uint32_t lane,
                                           const XchgArgs& xa)
{
    constexpr uint32_t kOut = MODE == kPospopcnt ? 16u : 32u;
    unsigned long long* mine = xa.buf[xa.rank];
    __threadfence();  // the other CTAs' atomics into acc happen-before their ticket
    const int world = xa.world;
    unsigned long long err = 0ull;
    if (world > 1) err = ld_relaxed_sys(mine + kXchgErr);  // in flight together with the exchange below
    unsigned long long v = 0ull;
    if (lane < kOut) v = atomicExch(mine + kXchgAcc + lane, 0ull);
    if (lane == 0) atomicExch(mine + kXchgTicket, 0ull);
    if (world <= 1) {
        if (lane < kOut) out[lane] = xa.accumulate ? out[lane] + v : v;
        return;
    }
    // 1. the epoch the previous launch left pending: its readers must be done with the slots
    //    of parity (epoch & 1) before step 2 overwrites them with epoch's own totals -- they are,
    //    because a peer pushes prev_epoch + 1 = epoch only after ITS step 1, and the acquire loads
    //    of step 1 keep this rank's own stores of step 2 behind its reads (a __threadfence_system()
    //    here cost 3 us per step: MEMBAR.SC.SYS on the critical path, profiles/r4x_*).
    // A failed exchange stays failed (the parity argument needs every rank to have completed
    // every epoch): nothing more is pushed or written, the host reports ESTATE / ETIMEOUT.
    if (xa.prev_epoch != 0ull) {
        if (!xchg_collect(xa.prev_out, lane, xa, xa.prev_epoch, xa.prev_accumulate, (uint32_t)xa.prev_nout, err))
            return;
    } else if (err) {
        if (lane == 0 && xa.host_err) st_relaxed_sys(xa.host_err, err);  // a peer's give-up reaches this host too
        return;
    }
    // 2. push this launch's totals to every rank: fire and forget
    const uint32_t par = (uint32_t)(xa.epoch & 1ull);
    const uint32_t slot = kXchgSlots + (par * kMaxRanks + (uint32_t)xa.rank) * kXchgSlotWords + 2u * lane;
    const unsigned long long tag = xchg_tag(xa.epoch);
    if (lane < kOut)
        for (int r = 0; r < world; ++r)
            st_relaxed_sys_v2(xa.buf[r] + slot, (v & 0xFFFFFFFFull) | tag, (v >> 32) | tag);
    // 3. unless deferred: wait for the peers' totals of this very epoch
    if (!xa.deferred) xchg_collect(out, lane, xa, xa.epoch, xa.accumulate, kOut);
}

// FLAGSTAT_cuda_xchg_collect: the pending epoch of a deferred launch, on its own (one warp)
__global__ void xchg_collect_kernel(const __grid_constant__ XchgArgs xa)
{
    if (xa.world > 1 && xa.prev_epoch != 0ull && ld_relaxed_sys(xa.buf[xa.rank] + kXchgErr) == 0ull)
        xchg_collect(xa.prev_out, threadIdx.x, xa, xa.prev_epoch, xa.prev_accumulate, (uint32_t)xa.prev_nout);
}

// Common tail of every kernel variant: CTA reduction of the per-warp position
// totals, mapping to counter slots, and either 64-bit atomics into out[] or
// the fused exchange above.
template <int MODE>
__device__ __forceinline__ void cta_epilogue(unsigned long long* __restrict__ out,
                                             unsigned long long acc_all, unsigned long long acc_fail,
                                             uint64_t n, uint32_t warp, uint32_t lane, bool counting_warp,
                                             const XchgArgs& xa)
{
    __shared__ unsigned long long s_all[kWarps][32];
    __shared__ unsigned long long s_fail[kWarps][32];
    if (counting_warp) {
        s_all[warp][lane] = acc_all;
        s_fail[warp][lane] = acc_fail;
    }
    __syncthreads();
    if (warp != 0) return;
    // nothing is published before the previous kernel of the stream has completed
    // (overlapped launches only)
    if (xa.pdl) pdl_wait_prior_grids();
    unsigned long long a = 0ull, f = 0ull;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
        a += s_all[i][lane];
        f += s_fail[i][lane];
    }
    // fold the high-halfword record onto the low one
    a += __shfl_down_sync(0xffffffffu, a, 16);
    f += __shfl_down_sync(0xffffffffu, f, 16);
    if (xa.world == 0) {
        emit_counters<MODE>(out, lane, a, f, n);
        return;
    }
    unsigned long long* mine = xa.buf[xa.rank];
    emit_counters<MODE>(mine + kXchgAcc, lane, a, f, n);
    __threadfence();
    __syncwarp();
    unsigned long long t = 0ull;
    if (lane == 0) t = atomicAdd(mine + kXchgTicket, 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t == (unsigned long long)gridDim.x - 1ull) xchg_last_cta<MODE>(out, lane, xa);
}

// Persistent, grid-strided over 16 KiB CTA batches.  out = uint64_t[32]
// (flagstat) or uint64_t[16] (pospopcnt), ACCUMULATED with 64-bit atomics.
//
// Work split: records [0, head) bring the base up to 16-byte alignment, V full
// 128-bit vectors follow, then < 8 tail records.  NB = V / kVecPerBatch full CTA
// batches are strided over the grid; the CTA that would own batch NB also takes
// the < kVecPerBatch left-over vectors (zero padded) and the <= 14 ragged
// records, as two extra batches in front of its main loop.
template <int MODE, int VARIANT>
__global__ void __launch_bounds__(kThreads, 2)
flagstat_kernel(const uint16_t* __restrict__ base, uint64_t n, unsigned long long* __restrict__ out,
                const __grid_constant__ XchgArgs xa)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t addr = reinterpret_cast<uint64_t>(base);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint4* __restrict__ body = reinterpret_cast<const uint4*>(base + head);
    const uint64_t V = (n - head) >> 3;
    const uint64_t tail_start = head + (V << 3);
    const uint64_t ntail = n - tail_start;
    const uint64_t NB = V / kVecPerBatch;
    const uint64_t G = gridDim.x;

    Lanes<MODE, VARIANT> st;
    st.clear();
    uint32_t b = 0;
    unsigned long long acc_all = 0ull, acc_fail = 0ull;  // lane j: total of position j

    if (blockIdx.x == (uint32_t)(NB % G)) {
        uint32_t w[16];
        {
            uint4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint64_t idx = NB * kVecPerBatch + (uint64_t)u * kThreads + tid;
                v[u] = (idx < V) ? ld_stream(body + idx) : make_uint4(0u, 0u, 0u, 0u);
            }
            unpack4(v, w);
        }
        st.step(w, b);
        ++b;
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
        if (tid < head) w[0] = base[tid];
        else if (tid - head < ntail) w[0] = base[tail_start + (tid - head)];
        st.step(w, b);
        ++b;
    }

    // This CTA's batches are blockIdx.x, blockIdx.x + G, ...: a 32-bit trip count
    // and one pointer bump per batch keep the loop bookkeeping off the ALU pipe.
    const uint32_t my = (NB > blockIdx.x) ? (uint32_t)((NB - blockIdx.x + G - 1) / G) : 0u;
    const uint64_t stride = G * (uint64_t)kVecPerBatch;
    const uint4* __restrict__ p = body + tid + (uint64_t)blockIdx.x * kVecPerBatch;
    uint4 bufA[kU], bufB[kU];
    uint32_t it = 0;
    if (my) load_batch(bufA, p);

    do {
        // one epoch: at most Counter::kMaxBatches batches, then expand
        for (;;) {
            if (it >= my || b >= Counter::kMaxBatches) break;
            {
                p += stride;
                if (it + 1 < my) load_batch(bufB, p);
                uint32_t w[16];
                unpack4(bufA, w);
                st.step(w, b);
                ++b;
                ++it;
            }
            if (it >= my) break;
            {
                p += stride;
                if (it + 1 < my) load_batch(bufA, p);
                uint32_t w[16];
                unpack4(bufB, w);
                st.step(w, b);
                ++b;
                ++it;
            }
        }
        acc_all += st.all.flush_warp(b, lane);
        if (MODE == kFlagstat && st.nfail != 0u) acc_fail += st.fail.flush_warp(st.nfail, lane);
        st.clear();
        b = 0;
    } while (it < my);

    cta_epilogue<MODE>(out, acc_all, acc_fail, n, warp, lane, true, xa);
}

// ---------------------------------------------------------------------------
// read-only HBM probe (FLAGSTAT_cuda_read_probe): the load side of a streaming
// kernel with the arithmetic taken out -- 8 LDG.128 in flight per thread, one XOR
// per 16 bytes.  The sink store never happens for real data.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
hbm_read_probe_kernel(const uint4* __restrict__ p, uint64_t nvec, unsigned long long* __restrict__ sink)
{
    constexpr int U = 8;
    uint32_t acc = 0u;
    const uint64_t per = (uint64_t)U * kThreads;
    const uint64_t nb = nvec / per;
    for (uint64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_stream(p + b * per + (uint64_t)u * kThreads + threadIdx.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (uint64_t i = nb * per + (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec;
         i += (uint64_t)gridDim.x * kThreads) {
        const uint4 v = ld_stream(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9E3779B9u && nvec == 1u) sink[0] = acc;  // keeps the loads alive
}

}  // namespace fsb200
