// zstd_frame.cuh -- Zstandard *frame* decoder for the reference's Zstd block files
// (benchmark/flagstats.cpp:192-215 writes [int32 raw_size][int32 comp_size][Zstd frame]
// records with ZSTD_compress, :90-93; zstd_decompress(), :636-676, decodes every record with
// ZSTD_decompress and feeds N = raw_size >> 1 records to the flagstat kernel).
//
// zstd is a system library the reference links; what is implemented here is the published
// format, RFC 8878: frame header, raw / RLE / compressed blocks, literals (raw, RLE, Huffman
// with 1 or 4 streams, treeless), Huffman tree descriptions (direct or FSE-compressed
// weights), sequences (predefined / RLE / FSE / repeat tables, backward bitstream), repeat
// offsets, sequence execution.  No dictionaries; the content checksum is skipped.
//
// Two ways to run it, one code path (the sequence loop hands every sequence to an EXECUTOR):
//   * decode_frame: ONE THREAD decodes one frame start to end, the executor copies at once
//     (first version; all state in a per-frame workspace in global memory);
//   * parse_frame (second version, what the library runs): entropy decoding only -- one lane
//     per frame, tables in shared memory, a 64-bit bit window in registers -- and the executor
//     RECORDS every sequence as a 16-byte descriptor {output position, literal position,
//     literal length, offset}; the copies are then done by a whole CTA per frame with the
//     machinery of the LZ4 decoder (lz4_block_cta.cuh, l4_copy: parents + pointer jumping).
// There is no allocation, no recursion and no unaligned multi-byte access.
// The whole decoder is FSB_HD (__host__ __device__): tests/test_zstd_frame_host.py compiles
// this very file with g++ and holds it to the real libzstd on the CPU; the GPU tests then only
// have to show that the same code gives the same bytes on the device.  The host instantiation
// exists for those tests only: libflagstats_cuda.so calls decode_frame from zstd_decode_kernel
// and nowhere else (there is no CPU fallback in the product).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FSB_HD __host__ __device__ __forceinline__
#define FSB_HDN __host__ __device__ __noinline__
#if defined(__CUDA_ARCH__)
#define FSB_UNROLL _Pragma("unroll")
#else
#define FSB_UNROLL  // (the host pass of nvcc is g++: it does not know the pragma)
#endif
#else
#define FSB_HD inline
#define FSB_HDN inline
#define FSB_UNROLL
#endif

namespace fsb200 {
namespace zstd {

constexpr int kErrTrunc = -1, kErrMagic = -2, kErrHeader = -3, kErrBlock = -4, kErrLiterals = -5, kErrHuf = -6,
              kErrFse = -7, kErrSeq = -8, kErrOut = -9;
constexpr uint32_t kBlockMax = 128u << 10;  // Block_Maximum_Size
constexpr int kFseMaxLog = 9, kHufMaxBits = 11;

struct FseTab {
    int log;  // accuracy log; -1 = no table yet
    uint8_t sym[1 << kFseMaxLog];
    uint8_t nbits[1 << kFseMaxLog];
    uint16_t base[1 << kFseMaxLog];
};
struct HufTab {
    int bits;  // 0 = no table yet
    uint8_t sym[1 << kHufMaxBits];
    uint8_t len[1 << kHufMaxBits];
};
// One state of a sequence table (LL / OF / ML), everything the sequence loop needs in one 8-byte load:
// bits 0-31 the base VALUE of the state's symbol (literal length, match length, 1 << offset code), 32-39
// the symbol's extra bits, 40-47 the bits to read for the next state, 48-63 the next state's base.
// Symbols are validated when the table is packed, not per sequence.
template <int MAXLOG>
struct SeqTab {
    int log;  // accuracy log; -1 = no table yet
    uint64_t st[1 << MAXLOG];
};
// per-frame tables (second version: shared memory, one set per warp)
struct Tables {
    SeqTab<9> ll, ml;  // accuracy logs <= 9 (LL, ML), <= 8 (OF): RFC 8878 3.1.1.3.2.1
    SeqTab<8> of;
    FseTab wt;  // FSE-compressed Huffman weights; also where a sequence table is built before it is packed
    HufTab huf;
    uint64_t rep[3];
    int16_t freq[256];
    uint16_t next[256];
    uint8_t w[260];
};
// per-frame workspace of the first version (global memory on the device)
struct Work {
    Tables t;
    uint8_t lit[kBlockMax + 32];
};
// what parse_frame writes per sequence (same layout as L4Desc of lz4_block_cta.cuh)
struct SeqDesc {
    uint32_t out_pos;  // first output byte of the sequence (its literals)
    uint32_t lit_pos;  // position of its first literal byte in the frame's literal buffer
    uint32_t lit;      // literal bytes; the match follows and ends where the next sequence starts
    uint32_t off;      // match offset (0: literals only)
};

FSB_HD int highbit(uint32_t v)  // floor(log2(v)), v > 0
{
    int h = 0;
    while (v >>= 1) ++h;
    return h;
}

// little-endian value of n <= 8 bytes, byte loads only
FSB_HD uint64_t le(const uint8_t* p, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// ---- copies ------------------------------------------------------------------------
//
// One thread copying byte by byte pays one memory round trip per byte on the device: the
// bytes a match reads were stored a moment ago (stores do not allocate in L1), and since
// source and destination may alias the compiler cannot overlap iterations.  Both copies below
// therefore move 16 bytes at a time, all loads first, then all stores: one round trip per 16
// bytes.  Plain byte accesses only (no alignment requirement, same code on the host).

// dst and src do not overlap, or dst - src >= 16
FSB_HD void copy16(uint8_t* dst, const uint8_t* src, uint64_t n)
{
    uint64_t k = 0;
    for (; k + 16 <= n; k += 16) {
        uint8_t t[16];
        FSB_UNROLL
        for (int j = 0; j < 16; ++j) t[j] = src[k + j];
        FSB_UNROLL
        for (int j = 0; j < 16; ++j) dst[k + j] = t[j];
    }
    for (; k < n; ++k) dst[k] = src[k];
}

// LZ77 match: dst[k] = dst[k - off] for k in [0, n), off >= 1, overlap allowed.  A match
// closer than 16 bytes repeats a pattern of period `off`; once the first (m - 1) * off bytes
// are out, copying from m * off bytes back gives the same bytes, so the distance is widened to
// >= 16 and the 16-byte copy applies.
FSB_HD void copy_match(uint8_t* dst, uint64_t off, uint64_t n)
{
    uint64_t k = 0;
    uint64_t dist = off;
    if (off < 16) {
        const uint64_t m = (15 + off) / off;  // smallest m with m * off >= 16
        dist = m * off;
        const uint64_t head = dist - off < n ? dist - off : n;
        for (; k < head; ++k) dst[k] = dst[k - off];
    }
    copy16(dst + k, dst + k - dist, n - k);
}

// ---- bit readers -------------------------------------------------------------------

// forward stream, bit 0 of byte 0 first (FSE table descriptions)
struct Fwd {
    const uint8_t* p;
    uint64_t nbits, pos;
    FSB_HD uint32_t read(int n)  // n <= 16
    {
        uint32_t v = 0;
        if (pos + 24 <= nbits) {
            const uint64_t b = pos >> 3;
            v = (uint32_t)((le(p + b, 3) >> (pos & 7)) & ((1u << n) - 1u));
        } else {
            for (int i = 0; i < n; ++i)
                if (pos + i < nbits) v |= (uint32_t)((p[(pos + i) >> 3] >> ((pos + i) & 7)) & 1u) << i;
        }
        pos += (uint64_t)n;
        return v;
    }
};

// backward stream: the last byte carries a 1 above the payload; reads walk towards byte 0
// and bits "before" the stream are zeros.  Bits [wlo, wlo + 64) of the stream are kept in a
// register window (eight byte loads per refill, every ~40 bits) so that the six reads of a
// sequence are a compare, a shift and a mask each; everything else (refill, the first bytes of
// the stream, reads that run off its start) is ONE out-of-line function, which keeps the
// sequence loop small enough to stay in the instruction cache with a frame per warp.
struct Back {
    const uint8_t* p;
    int32_t pos;
    int32_t wlo;   // bit index of the window's bit 0 (a multiple of 8); the window is empty if wlo > pos
    uint64_t win;
    FSB_HD bool init(const uint8_t* q, uint64_t n)
    {
        if (n == 0 || n > ((uint64_t)1 << 27) || q[n - 1] == 0) return false;  // (streams live inside 128 KiB blocks)
        p = q;
        pos = (int32_t)(n - 1) * 8 + highbit(q[n - 1]);
        wlo = pos + 1;  // nothing loaded yet
        win = 0;
        return true;
    }
    FSB_HD uint32_t read(int n)  // n <= 32
    {
        pos -= n;
        // (reads only ever move down: pos + n <= wlo + 64 holds since the refill)
        const uint32_t mask = (uint32_t)(((uint64_t)1 << n) - 1u);
        if (pos >= wlo) return (uint32_t)(win >> (pos - wlo)) & mask;
        if (pos >= 56) {  // the usual refill, in line: the window ends with the byte that holds bit pos + n - 1
            const int32_t top = (pos + (n ? n - 1 : 0)) >> 3;  // >= 7
            wlo = (top - 7) * 8;
            const uint8_t* q = p + (top - 7);
            win = (uint64_t)q[0] | ((uint64_t)q[1] << 8) | ((uint64_t)q[2] << 16) | ((uint64_t)q[3] << 24) |
                  ((uint64_t)q[4] << 32) | ((uint64_t)q[5] << 40) | ((uint64_t)q[6] << 48) | ((uint64_t)q[7] << 56);
            return (uint32_t)(win >> (pos - wlo)) & mask;
        }
        return slow(p, pos, n);
    }
    // (a static function of VALUES: a member taking `this` would force the reader out of its registers and
    // into local memory -- measured: every read then goes through the stack, ~100 cycles each)
    FSB_HDN static uint32_t slow(const uint8_t* p, int32_t pos, int n)  // pos has been moved already
    {
        if (n == 0) return 0u;
        const uint32_t mask = (uint32_t)(((uint64_t)1 << n) - 1u);
        if (pos >= 0) {
            // within the first bytes of the stream (pos < 56): bits [pos, pos + n) lie inside at most 5 bytes
            const int32_t top = (pos + n - 1) >> 3;  // last byte needed
            const int32_t b = pos >> 3;
            const int nb = (int)(top - b) + 1;
            return (uint32_t)(le(p + b, nb) >> (pos & 7)) & mask;
        }
        uint32_t v = 0;
        for (int i = 0; i < n; ++i) {
            const int32_t q = pos + i;
            if (q >= 0) v |= (uint32_t)((p[q >> 3] >> (q & 7)) & 1u) << i;
        }
        return v;
    }
};

// ---- FSE (RFC 8878 4.1) ---------------------------------------------------------------

FSB_HDN int fse_build(FseTab& t, const int16_t* freq, int nsym, int log, uint16_t* next)
{
    const int size = 1 << log;
    int high = size;
    if (log > kFseMaxLog || nsym > 256) return kErrFse;
    for (int s = 0; s < nsym; ++s)
        if (freq[s] == -1) {
            t.sym[--high] = (uint8_t)s;
            next[s] = 1;
        }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; ++s) {
        if (freq[s] <= 0) continue;
        next[s] = (uint16_t)freq[s];
        for (int i = 0; i < freq[s]; ++i) {
            t.sym[pos] = (uint8_t)s;
            do pos = (pos + step) & mask; while (pos >= high);
        }
    }
    if (pos != 0) return kErrFse;
    for (int i = 0; i < size; ++i) {
        const uint16_t n = next[t.sym[i]]++;
        t.nbits[i] = (uint8_t)(log - highbit(n));
        t.base[i] = (uint16_t)(((uint32_t)n << t.nbits[i]) - (uint32_t)size);
    }
    t.log = log;
    return 0;
}

// table description -> table; returns bytes consumed or < 0
FSB_HDN int64_t fse_read(FseTab& t, const uint8_t* p, uint64_t n, int max_log, int max_sym, Tables& w)
{
    Fwd b{p, n * 8, 0};
    const int log = 5 + (int)b.read(4);
    if (log > max_log) return kErrFse;
    int remaining = 1 << log, s = 0;
    while (remaining > 0 && s <= max_sym) {
        const int bits = highbit((uint32_t)remaining + 1u) + 1;
        uint32_t v = b.read(bits);
        const uint32_t lower = (1u << (bits - 1)) - 1u;
        const uint32_t thresh = (1u << bits) - 1u - ((uint32_t)remaining + 1u);
        if ((v & lower) < thresh) {
            b.pos -= 1;
            v &= lower;
        } else if (v > lower) {
            v -= thresh;
        }
        const int proba = (int)v - 1;
        remaining -= proba < 0 ? 1 : proba;
        w.freq[s++] = (int16_t)proba;
        if (proba == 0) {
            for (;;) {
                const int rep = (int)b.read(2);
                for (int i = 0; i < rep && s <= max_sym; ++i) w.freq[s++] = 0;
                if (rep != 3) break;
            }
        }
        if (b.pos > b.nbits) return kErrTrunc;
    }
    if (remaining != 0 || s > max_sym + 1) return kErrFse;
    const int rc = fse_build(t, w.freq, s, log, w.next);
    if (rc) return rc;
    return (int64_t)((b.pos + 7) >> 3);
}

// ---- Huffman (RFC 8878 4.2) --------------------------------------------------------------

FSB_HDN int huf_build(HufTab& h, uint8_t* wt, int n)  // wt[0..n-1] given, wt[n] implied
{
    uint32_t sum = 0;
    for (int i = 0; i < n; ++i) {
        if (wt[i] > kHufMaxBits) return kErrHuf;
        if (wt[i]) sum += 1u << (wt[i] - 1);
    }
    if (sum == 0 || n >= 256) return kErrHuf;
    const int max_bits = highbit(sum) + 1;
    if (max_bits > kHufMaxBits) return kErrHuf;
    const uint32_t left = (1u << max_bits) - sum;
    if ((left & (left - 1u)) != 0u) return kErrHuf;
    wt[n++] = (uint8_t)(highbit(left) + 1);
    uint32_t start[kHufMaxBits + 2];
    for (int k = 0; k < kHufMaxBits + 2; ++k) start[k] = 0;
    for (int i = 0; i < n; ++i)
        if (wt[i]) start[wt[i] + 1] += 1u << (wt[i] - 1);
    for (int k = 1; k < kHufMaxBits + 2; ++k) start[k] += start[k - 1];
    for (int i = 0; i < n; ++i) {  // by increasing weight, then by symbol value
        if (!wt[i]) continue;
        const uint32_t span = 1u << (wt[i] - 1);
        const uint8_t len = (uint8_t)(max_bits + 1 - wt[i]);
        const uint32_t at = start[wt[i]];
        for (uint32_t k = 0; k < span; ++k) {
            h.sym[at + k] = (uint8_t)i;
            h.len[at + k] = len;
        }
        start[wt[i]] = at + span;
    }
    h.bits = max_bits;
    return 0;
}

// tree description; returns bytes consumed or < 0
FSB_HDN int64_t huf_read_tree(Tables& w, const uint8_t* p, uint64_t n)
{
    if (n < 1) return kErrTrunc;
    const int hb = p[0];
    if (hb >= 128) {  // direct: 4 bits per weight
        const int cnt = hb - 127;
        const uint64_t bytes = ((uint64_t)cnt + 1) / 2;
        if (n < 1 + bytes) return kErrTrunc;
        for (int i = 0; i < cnt; ++i) w.w[i] = (i & 1) ? (uint8_t)(p[1 + i / 2] & 15) : (uint8_t)(p[1 + i / 2] >> 4);
        const int rc = huf_build(w.huf, w.w, cnt);
        return rc ? rc : (int64_t)(1 + bytes);
    }
    if (hb == 0 || n < 1 + (uint64_t)hb) return kErrTrunc;
    const int64_t used = fse_read(w.wt, p + 1, (uint64_t)hb, 6, 255, w);
    if (used < 0) return used;
    if (used >= hb) return kErrHuf;
    Back b;
    if (!b.init(p + 1 + used, (uint64_t)hb - (uint64_t)used)) return kErrHuf;
    const FseTab& t = w.wt;
    uint32_t s1 = b.read(t.log), s2 = b.read(t.log);
    if (b.pos < 0) return kErrHuf;
    int cnt = 0;
    for (;;) {  // two interleaved states until the stream runs dry
        if (cnt >= 254) return kErrHuf;
        w.w[cnt++] = t.sym[s1];
        s1 = t.base[s1] + b.read(t.nbits[s1]);
        if (b.pos < 0) {
            w.w[cnt++] = t.sym[s2];
            break;
        }
        if (cnt >= 254) return kErrHuf;
        w.w[cnt++] = t.sym[s2];
        s2 = t.base[s2] + b.read(t.nbits[s2]);
        if (b.pos < 0) {
            w.w[cnt++] = t.sym[s1];
            break;
        }
    }
    const int rc = huf_build(w.huf, w.w, cnt);
    return rc ? rc : (int64_t)(1 + hb);
}

// one Huffman stream -> exactly `want` symbols
FSB_HDN int huf_stream(const HufTab& h, const uint8_t* p, uint64_t n, uint8_t* out, uint64_t want)
{
    Back b;
    if (!b.init(p, n)) return kErrHuf;
    const uint32_t mask = (1u << h.bits) - 1u;
    uint32_t state = b.read(h.bits);
    uint64_t got = 0;
    while (b.pos > -(int64_t)h.bits) {
        if (got == want) return kErrHuf;
        out[got++] = h.sym[state];
        const int l = h.len[state];
        state = ((state << l) + b.read(l)) & mask;
    }
    if (b.pos != -(int64_t)h.bits || got != want) return kErrHuf;
    return 0;
}

// ---- sequences (RFC 8878 3.1.1.3.2) ----------------------------------------------------------

FSB_HD uint32_t ll_base(int c) { return c < 16 ? (uint32_t)c : c < 20 ? 16u + 2u * (uint32_t)(c - 16) : c < 22 ? 24u + 4u * (uint32_t)(c - 20) : c < 24 ? 32u + 8u * (uint32_t)(c - 22) : c == 24 ? 48u : 64u << (c - 25); }
FSB_HD int ll_bits(int c) { return c < 16 ? 0 : c < 20 ? 1 : c < 22 ? 2 : c < 24 ? 3 : c == 24 ? 4 : c - 19; }
FSB_HD uint32_t ml_base(int c)
{
    return c < 32 ? 3u + (uint32_t)c : c < 36 ? 35u + 2u * (uint32_t)(c - 32) : c < 38 ? 43u + 4u * (uint32_t)(c - 36)
         : c < 40 ? 51u + 8u * (uint32_t)(c - 38) : c < 42 ? 67u + 16u * (uint32_t)(c - 40) : c == 42 ? 99u
         : (128u << (c - 43)) + 3u;
}
FSB_HD int ml_bits(int c) { return c < 32 ? 0 : c < 36 ? 1 : c < 38 ? 2 : c < 40 ? 3 : c < 42 ? 4 : c == 42 ? 5 : c - 36; }

// predefined distributions, RFC 8878 3.1.1.3.2.2 (written as code so that no static table has to
// live in device constant memory: -1 entries are the "less than 1" probabilities)
FSB_HD int16_t ll_default(int s)
{
    return s == 0 ? 4 : s == 1 ? 3 : s <= 12 ? 2 : s <= 15 ? 1 : s <= 24 ? 2 : s == 25 ? 3 : s == 26 ? 2 : s <= 31 ? 1 : -1;
}
FSB_HD int16_t ml_default(int s) { return s == 0 ? 1 : s == 1 ? 4 : s == 2 ? 3 : s <= 8 ? 2 : s <= 45 ? 1 : -1; }
FSB_HD int16_t of_default(int s) { return s <= 5 ? 1 : s <= 8 ? 2 : s <= 23 ? 1 : -1; }

// one of the three tables of a sequences section; which: 0 = LL, 1 = OF, 2 = ML.  The FSE table is
// built in w.wt (free at this point: the literals are done) and packed into t.  Returns bytes consumed or < 0.
template <class Tab>
FSB_HDN int64_t seq_table(Tab& t, int mode, const uint8_t* p, uint64_t n, int which, Tables& w)
{
    const int def_n = which == 0 ? 36 : which == 1 ? 29 : 53;
    const int def_log = which == 1 ? 5 : 6;
    const int max_log = which == 1 ? 8 : 9;
    const int max_sym = which == 0 ? 35 : which == 1 ? 31 : 52;
    int64_t used = 0;
    FseTab& f = w.wt;
    if (mode == 0) {
        for (int s = 0; s < def_n; ++s) w.freq[s] = which == 0 ? ll_default(s) : which == 1 ? of_default(s) : ml_default(s);
        const int rc = fse_build(f, w.freq, def_n, def_log, w.next);
        if (rc) return rc;
    } else if (mode == 1) {
        if (n < 1) return kErrTrunc;
        if (p[0] > max_sym) return kErrSeq;
        f.log = 0;
        f.sym[0] = p[0];
        f.nbits[0] = 0;
        f.base[0] = 0;
        used = 1;
    } else if (mode == 2) {
        used = fse_read(f, p, n, max_log, max_sym, w);
        if (used < 0) return used;
    } else {
        return t.log < 0 ? kErrSeq : 0;  // repeat: the previous table must exist
    }
    const int size = 1 << f.log;
    for (int i = 0; i < size; ++i) {
        const int c = f.sym[i];
        if (c > max_sym) return kErrSeq;
        const uint32_t value = which == 0 ? ll_base(c) : which == 1 ? (uint32_t)1 << c : ml_base(c);
        const uint32_t extra = (uint32_t)(which == 0 ? ll_bits(c) : which == 1 ? c : ml_bits(c));
        t.st[i] = (uint64_t)value | ((uint64_t)extra << 32) | ((uint64_t)f.nbits[i] << 40) | ((uint64_t)f.base[i] << 48);
    }
    t.log = f.log;
    return used;
}

// ---- executors: what happens to a decoded sequence -----------------------------------------

// first version: copy at once into out[op..]
struct ExecCopy {
    static constexpr bool kKeepLiterals = false;  // literals may be used where they lie (input, block buffer)
    static constexpr bool kDeferChecks = false;   // it writes the output: every sequence is checked before it is executed
    uint8_t* out;
    uint64_t op, cap;
    uint32_t bad = 0;  // (unused: interface of the sequence loop)
    // ll literal bytes at lit, then a match of ml bytes from `off` back (ml == 0: literals only)
    FSB_HD int seq(uint32_t ll, uint32_t ml, uint32_t off, const uint8_t* lit, uint32_t /*lit_pos*/)
    {
        if (ll > cap - op || ml > cap - op - ll) return kErrOut;
        copy16(out + op, lit, ll);
        op += ll;
        if (ml) {
            if (off == 0 || off > op) return kErrSeq;
            copy_match(out + op, off, ml);
            op += ml;
        }
        return 0;
    }
    FSB_HD int check() const { return 0; }
    FSB_HD int fill(uint8_t v, uint64_t n, uint8_t* /*litbuf*/, uint32_t /*lit_pos*/)  // an RLE block
    {
        if (n > cap - op) return kErrOut;
        for (uint64_t k = 0; k < n; ++k) out[op + k] = v;
        op += n;
        return 0;
    }
};

// second version: record the sequence; the copies are another kernel's business.  Every literal byte
// of the frame ends up in ONE literal buffer (a literal byte is an output byte, so raw-size bytes
// are enough), which is what the descriptors index.
struct ExecRecord {
    static constexpr bool kKeepLiterals = true;
    // It only WRITES DESCRIPTORS (bounded by cap_d), so nothing a corrupt sequence says can reach memory: the
    // sequence loop runs without early exits -- violations are OR-ed into `bad` and looked at once per block.
    // One lane runs that loop; a compare-and-branch per check was a fifth of its instructions and most of its
    // branch stalls.
    static constexpr bool kDeferChecks = true;
    SeqDesc* d;
    uint32_t nd, cap_d;
    uint32_t op, cap;  // (32-bit on purpose: this runs once per sequence on ONE lane, every instruction is latency)
    uint32_t bad = 0;  // sticky: a sequence overran the output, pointed in front of it, or d[] is full
    // the last descriptor, if it is literals only: kept HERE, not read back from d[] (on the device that
    // read would be a global-memory round trip per sequence)
    bool open = false;
    uint32_t open_lit = 0;      // its literal bytes so far
    uint32_t open_lit_end = 0;  // position behind them in the literal buffer
    FSB_HD int seq(uint32_t ll, uint32_t ml, uint32_t off, const uint8_t* /*lit*/, uint32_t lit_pos)
    {
        const uint64_t m = (uint64_t)op + ll, end = m + ml;
        bad |= (uint32_t)(end > cap) | (uint32_t)(ml != 0u && (off == 0u || off > m));
        if ((ll | ml) == 0u) return 0;
        if (open && open_lit_end == lit_pos) {
            // the descriptor before was literals only (a block's tail, a raw block; it ends at op by
            // construction) and these literals follow them in the buffer: one descriptor.  Every descriptor
            // but the last then carries a match of >= 3 bytes, so cap / 3 + 2 descriptors are always enough.
            open_lit += ll;
            d[nd - 1u].lit = open_lit;
            d[nd - 1u].off = ml ? off : 0u;
        } else if (nd < cap_d) {
            d[nd++] = SeqDesc{op, lit_pos, ll, ml ? off : 0u};
            open_lit = ll;
        } else {
            bad |= 1u;
            return 0;  // (`open` keeps implying that d[nd - 1] exists)
        }
        open = ml == 0u;
        open_lit_end = lit_pos + ll;
        op = (uint32_t)end;  // (meaningless once bad is set; nothing is addressed with it)
        return 0;
    }
    FSB_HD int check() const { return bad ? kErrSeq : 0; }
    FSB_HD int fill(uint8_t v, uint64_t n, uint8_t* litbuf, uint32_t lit_pos)  // an RLE block: one literal + a run
    {
        if (n == 0) return 0;
        if (n > cap - op) return kErrOut;
        litbuf[0] = v;
        return seq(1u, (uint32_t)n - 1u, 1u, litbuf, lit_pos);
    }
};

// a compressed block: literals + sequences -> the executor; returns the literal bytes the block
// regenerated (they start at litbuf[0] = literal position lit_pos of the frame) or < 0.
// litbuf has room for kBlockMax + 32 bytes.
template <class Exec>
FSB_HDN int64_t block_compressed(Tables& c, uint8_t* litbuf, uint32_t lit_pos, const uint8_t* p, uint64_t n, Exec& ex)
{
    // ---- literals section ----
    if (n < 1) return kErrTrunc;
    const int ltype = p[0] & 3, sf = (p[0] >> 2) & 3;
    uint64_t regen, comp = 0, hdr;
    const uint8_t* lit;
    if (ltype < 2) {  // raw / RLE
        if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return kErrTrunc; hdr = 2; regen = le(p, 2) >> 4; }
        else { if (n < 3) return kErrTrunc; hdr = 3; regen = le(p, 3) >> 4; }
        if (regen > kBlockMax) return kErrLiterals;
        if (ltype == 0) {
            if (n < hdr + regen) return kErrTrunc;
            if (Exec::kKeepLiterals) {
                copy16(litbuf, p + hdr, regen);
                lit = litbuf;
            } else {
                lit = p + hdr;
            }
            comp = regen;
        } else {
            if (n < hdr + 1) return kErrTrunc;
            const uint8_t v = p[hdr];
            for (uint64_t i = 0; i < regen; ++i) litbuf[i] = v;
            lit = litbuf;
            comp = 1;
        }
    } else {  // Huffman-compressed / treeless
        int streams;
        if (sf == 0 || sf == 1) {
            if (n < 3) return kErrTrunc;
            hdr = 3;
            const uint64_t v = le(p, 3);
            regen = (v >> 4) & 0x3FF;
            comp = (v >> 14) & 0x3FF;
            streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return kErrTrunc;
            hdr = 4;
            const uint64_t v = le(p, 4);
            regen = (v >> 4) & 0x3FFF;
            comp = (v >> 18) & 0x3FFF;
            streams = 4;
        } else {
            if (n < 5) return kErrTrunc;
            hdr = 5;
            const uint64_t v = le(p, 5);
            regen = (v >> 4) & 0x3FFFF;
            comp = (v >> 22) & 0x3FFFF;
            streams = 4;
        }
        if (regen > kBlockMax || n < hdr + comp) return kErrLiterals;
        const uint8_t* q = p + hdr;
        uint64_t left = comp;
        if (ltype == 2) {
            const int64_t used = huf_read_tree(c, q, left);
            if (used < 0) return used;
            q += used;
            left -= (uint64_t)used;
        } else if (c.huf.bits == 0) {
            return kErrHuf;  // treeless without a previous tree
        }
        if (streams == 1) {
            const int rc = huf_stream(c.huf, q, left, litbuf, regen);
            if (rc) return rc;
        } else {
            if (left < 6) return kErrLiterals;
            const uint64_t s1 = le(q, 2), s2 = le(q + 2, 2), s3 = le(q + 4, 2);
            if (6 + s1 + s2 + s3 > left) return kErrLiterals;
            const uint64_t s4 = left - 6 - s1 - s2 - s3;
            const uint64_t each = (regen + 3) / 4;
            if (3 * each > regen) return kErrLiterals;
            int rc;
            if ((rc = huf_stream(c.huf, q + 6, s1, litbuf, each))) return rc;
            if ((rc = huf_stream(c.huf, q + 6 + s1, s2, litbuf + each, each))) return rc;
            if ((rc = huf_stream(c.huf, q + 6 + s1 + s2, s3, litbuf + 2 * each, each))) return rc;
            if ((rc = huf_stream(c.huf, q + 6 + s1 + s2 + s3, s4, litbuf + 3 * each, regen - 3 * each))) return rc;
        }
        lit = litbuf;
    }
    p += hdr + comp;
    n -= hdr + comp;

    // ---- sequences section ----
    if (n < 1) return kErrTrunc;
    uint64_t nseq;
    if (p[0] == 0) { nseq = 0; p += 1; n -= 1; }
    else if (p[0] < 128) { nseq = p[0]; p += 1; n -= 1; }
    else if (p[0] < 255) { if (n < 2) return kErrTrunc; nseq = ((uint64_t)(p[0] - 128) << 8) + p[1]; p += 2; n -= 2; }
    else { if (n < 3) return kErrTrunc; nseq = (uint64_t)p[1] + ((uint64_t)p[2] << 8) + 0x7F00; p += 3; n -= 3; }
    uint32_t lp = 0;  // literals consumed
    const uint32_t nlit = (uint32_t)regen;
    if (nseq) {
        if (n < 1) return kErrTrunc;
        const int modes = p[0];
        if (modes & 3) return kErrSeq;
        p += 1; n -= 1;
        int64_t used;
        if ((used = seq_table(c.ll, modes >> 6, p, n, 0, c)) < 0) return used;
        p += used; n -= (uint64_t)used;
        if ((used = seq_table(c.of, (modes >> 4) & 3, p, n, 1, c)) < 0) return used;
        p += used; n -= (uint64_t)used;
        if ((used = seq_table(c.ml, (modes >> 2) & 3, p, n, 2, c)) < 0) return used;
        p += used; n -= (uint64_t)used;
        Back b;
        if (!b.init(p, n)) return kErrSeq;
        uint32_t sl = b.read(c.ll.log), so = b.read(c.of.log), sm = b.read(c.ml.log);
        // 32-bit arithmetic throughout the loop: one lane runs it, every instruction is a full pipeline latency
        uint32_t rep0 = (uint32_t)c.rep[0], rep1 = (uint32_t)c.rep[1], rep2 = (uint32_t)c.rep[2];
        const uint32_t ns = (uint32_t)nseq;
        Exec e = ex;  // (a copy the loop keeps in registers; `ex` itself lives in the caller's frame)
        for (uint32_t i = 0; i < ns; ++i) {
            const uint64_t eo = c.of.st[so], em = c.ml.st[sm], el = c.ll.st[sl];
            const uint32_t eoh = (uint32_t)(eo >> 32), emh = (uint32_t)(em >> 32), elh = (uint32_t)(el >> 32);
            const bool more = i + 1u < ns;
            // the six bit fields of a sequence, in stream order: offset, match length, literal length extra
            // bits, then the bits of the next LL, ML, OF states (none behind the last sequence)
            const int n1 = (int)(eoh & 0xFFu), n2 = (int)(emh & 0xFFu), n3 = (int)(elh & 0xFFu);
            const int n4 = more ? (int)((elh >> 8) & 0xFFu) : 0, n5 = more ? (int)((emh >> 8) & 0xFFu) : 0,
                      n6 = more ? (int)((eoh >> 8) & 0xFFu) : 0;
            const int total = n1 + n2 + n3 + n4 + n5 + n6;
            uint32_t ov, ml, ll;
            if (total <= 57 && b.pos >= 57) {
                // All six at once: the counts are known from the table entries, so the six positions are a
                // prefix sum and the six extractions are independent of one another -- one lane runs this
                // loop, and what it waits for is the LENGTH of the dependent chain, not the instruction count.
                if (b.pos - total < b.wlo) {  // window: the 64 bits that end with the byte holding bit pos - 1
                    const int32_t top = (b.pos - 1) >> 3;  // >= 7
                    const uint8_t* q = b.p + (top - 7);
                    b.wlo = (top - 7) * 8;  // <= pos - 57 <= pos - total
                    b.win = (uint64_t)q[0] | ((uint64_t)q[1] << 8) | ((uint64_t)q[2] << 16) | ((uint64_t)q[3] << 24) |
                            ((uint64_t)q[4] << 32) | ((uint64_t)q[5] << 40) | ((uint64_t)q[6] << 48) | ((uint64_t)q[7] << 56);
                }
                const int32_t base = b.pos - b.wlo;
                const int p1 = base - n1, p2 = p1 - n2, p3 = p2 - n3, p4 = p3 - n4, p5 = p4 - n5, p6 = p5 - n6;
                const uint64_t w = b.win;  // (a position is 64 only in front of an empty field: masked to 0 bits)
                ov = (uint32_t)eo + ((uint32_t)(w >> (p1 & 63)) & ((1u << n1) - 1u));  // (every count <= 31 here)
                ml = (uint32_t)em + ((uint32_t)(w >> (p2 & 63)) & ((1u << n2) - 1u));
                ll = (uint32_t)el + ((uint32_t)(w >> (p3 & 63)) & ((1u << n3) - 1u));
                if (more) {
                    sl = (elh >> 16) + ((uint32_t)(w >> (p4 & 63)) & ((1u << n4) - 1u));
                    sm = (emh >> 16) + ((uint32_t)(w >> (p5 & 63)) & ((1u << n5) - 1u));
                    so = (eoh >> 16) + ((uint32_t)(w >> (p6 & 63)) & ((1u << n6) - 1u));
                }
                b.pos -= total;
            } else {  // long offset codes, the first bytes of the stream, streams that run dry: one read at a time
                ov = (uint32_t)eo + b.read(n1);  // (1 << code) + code bits: < 2^32
                ml = (uint32_t)em + b.read(n2);
                ll = (uint32_t)el + b.read(n3);
                if (more) {
                    sl = (elh >> 16) + b.read(n4);
                    sm = (emh >> 16) + b.read(n5);
                    so = (eoh >> 16) + b.read(n6);
                }
            }
            if (Exec::kDeferChecks) e.bad |= (uint32_t)(b.pos < 0);  // (reads in front of the stream give zeros)
            else if (b.pos < 0) return kErrSeq;
            // repeat offsets, RFC 8878 3.1.1.5
            uint32_t off;
            if (ov > 3u) {
                off = ov - 3u;
                rep2 = rep1; rep1 = rep0; rep0 = off;
            } else {
                const uint32_t idx = ov - 1u + (ll == 0u ? 1u : 0u);
                if (idx == 0u) {
                    off = rep0;
                } else {
                    off = idx == 1u ? rep1 : idx == 2u ? rep2 : rep0 - 1u;
                    if (idx > 1u) rep2 = rep1;
                    rep1 = rep0;
                    rep0 = off;
                }
            }
            // execute: literals, then the match (which may overlap its own output)
            if (Exec::kDeferChecks) {
                e.bad |= (uint32_t)(ll > nlit - lp) | (uint32_t)(off == 0u);
                e.seq(ll, ml, off, lit + lp, lit_pos + lp);  // (lit + lp is not dereferenced by this executor)
            } else {
                if (ll > nlit - lp) return kErrOut;
                if (off == 0u) return kErrSeq;
                const int rc = e.seq(ll, ml, off, lit + lp, lit_pos + lp);
                if (rc) return rc;
            }
            lp += ll;
        }
        ex = e;
        if (ex.check()) return ex.check();
        c.rep[0] = rep0; c.rep[1] = rep1; c.rep[2] = rep2;
        if (b.pos != 0) return kErrSeq;
    }
    {
        const int rc = ex.seq(nlit - lp, 0u, 0u, lit + lp, lit_pos + lp);
        if (rc) return rc;
        if (ex.check()) return ex.check();
    }
    return (int64_t)regen;
}

// One frame through an executor; returns the output bytes (ex.op) or < 0.  c: tables, contents on
// entry do not matter.  lit / lit_cap: literal buffer -- for ExecCopy one block's worth
// (kBlockMax + 32, reused by every block), for ExecRecord the whole frame's (>= cap + 32: every
// literal byte is an output byte).
template <class Exec>
FSB_HDN int64_t run_frame(const uint8_t* in, uint64_t n, uint64_t cap, Tables& c, uint8_t* lit, uint64_t lit_cap, Exec& ex,
                          uint64_t* lit_used)
{
    if (n < 6) return kErrTrunc;
    if (le(in, 4) != 0xFD2FB528u) return kErrMagic;
    const int fhd = in[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, dict = fhd & 3;
    if (fhd & 0x08) return kErrHeader;
    uint64_t ip = 5;
    if (!single) ip += 1;  // window descriptor: the whole frame is decoded into one buffer anyway
    const int dict_bytes = dict == 3 ? 4 : dict;
    if (dict_bytes) {
        if (n < ip + (uint64_t)dict_bytes) return kErrTrunc;
        if (le(in + ip, dict_bytes) != 0) return kErrHeader;  // dictionaries are not supported
        ip += (uint64_t)dict_bytes;
    }
    const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (n < ip + (uint64_t)fcs_bytes) return kErrTrunc;
    uint64_t fcs = le(in + ip, fcs_bytes);
    if (fcs_bytes == 2) fcs += 256;
    ip += (uint64_t)fcs_bytes;
    if (fcs_bytes && fcs > cap) return kErrOut;

    c.ll.log = c.of.log = c.ml.log = c.wt.log = -1;
    c.huf.bits = 0;
    c.rep[0] = 1; c.rep[1] = 4; c.rep[2] = 8;
    uint64_t lpos = 0;  // literal bytes kept so far (ExecRecord); ExecCopy reuses the buffer
    for (;;) {
        if (n < ip + 3) return kErrTrunc;
        const uint32_t bh = (uint32_t)le(in + ip, 3);
        ip += 3;
        const int last = (int)(bh & 1u), type = (int)((bh >> 1) & 3u);
        const uint64_t size = bh >> 3;
        if (Exec::kKeepLiterals && lit_cap - lpos < kBlockMax + 32u) return kErrOut;  // (cannot happen with lit_cap >= cap + block)
        uint8_t* lb = Exec::kKeepLiterals ? lit + lpos : lit;
        if (type == 0) {
            if (n < ip + size) return kErrTrunc;
            if (size > cap - ex.op) return kErrOut;  // (hence size <= lit_cap - lpos: a literal byte is an output byte)
            const uint8_t* src = in + ip;
            if (Exec::kKeepLiterals) {
                copy16(lb, src, size);
                src = lb;
            }
            const int rc = ex.seq((uint32_t)size, 0u, 0u, src, (uint32_t)lpos);
            if (rc) return rc;
            if (Exec::kKeepLiterals) lpos += size;
            ip += size;
        } else if (type == 1) {
            if (n < ip + 1) return kErrTrunc;
            const int rc = ex.fill(in[ip], size, lb, (uint32_t)lpos);
            if (rc) return rc;
            if (Exec::kKeepLiterals && size) lpos += 1;
            ip += 1;
        } else if (type == 2) {
            if (size > kBlockMax) return kErrBlock;
            if (n < ip + size) return kErrTrunc;
            const int64_t r = block_compressed(c, lb, (uint32_t)lpos, in + ip, size, ex);
            if (r < 0) return r;
            if (Exec::kKeepLiterals) lpos += (uint64_t)r;
            ip += size;
        } else {
            return kErrBlock;
        }
        if (last) break;
    }
    if (checksum) {
        if (n < ip + 4) return kErrTrunc;
        ip += 4;  // xxh64 of the content, low 32 bits: not verified
    }
    if (ex.check()) return ex.check();
    if (fcs_bytes && fcs != ex.op) return kErrOut;
    if (lit_used) *lit_used = lpos;
    return (int64_t)ex.op;
}

// First version.  One frame; returns bytes produced or < 0.  `w` is scratch, its contents on entry do not matter.
FSB_HDN int64_t decode_frame(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, Work& w)
{
    ExecCopy ex{out, 0, cap};
    return run_frame(in, n, cap, w.t, w.lit, sizeof(w.lit), ex, nullptr);
}

// Second version, entropy stage.  One frame -> descriptors d[0, *nd) (at most cap_d) and the frame's
// literal bytes lit[0, *lit_used) (lit_cap >= cap + kBlockMax + 64); returns the bytes the descriptors
// produce or < 0.
FSB_HDN int64_t parse_frame(const uint8_t* in, uint64_t n, uint64_t cap, Tables& t, uint8_t* lit, uint64_t lit_cap,
                            SeqDesc* d, uint32_t cap_d, uint32_t* nd, uint64_t* lit_used)
{
    if (cap > 0xFFFFFFFFull) return kErrOut;
    ExecRecord ex;
    ex.d = d; ex.nd = 0u; ex.cap_d = cap_d; ex.op = 0u; ex.cap = (uint32_t)cap;
    const int64_t r = run_frame(in, n, cap, t, lit, lit_cap, ex, lit_used);
    *nd = ex.nd;
    return r;
}

// How many descriptors parse_frame can write for this frame, from the block headers alone (no entropy
// decoding: frame header, block headers, literals-section and sequences-section headers): the sum of the
// blocks' sequence counts + one literals-only descriptor per block + 2.  The library sizes each frame's slice
// of the descriptor scratch with it (the worst case, cap / 3 + 2, is 5.5 MB for a 1,024,000-byte frame of
// which FLAG data uses a seventh).  Anything odd in the headers: the worst case, and parse_frame reports it.
FSB_HDN uint64_t count_descriptors(const uint8_t* in, uint64_t n, uint64_t cap)
{
    const uint64_t worst = cap / 3u + 2u;
    if (n < 6 || le(in, 4) != 0xFD2FB528u) return worst;
    const int fhd = in[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, dict = fhd & 3;
    uint64_t ip = 5 + (single ? 0 : 1) + (uint64_t)(dict == 3 ? 4 : dict);
    ip += (uint64_t)(fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8);
    uint64_t total = 2;
    for (;;) {
        if (n < ip + 3) return worst;
        const uint32_t bh = (uint32_t)le(in + ip, 3);
        ip += 3;
        const int last = (int)(bh & 1u), type = (int)((bh >> 1) & 3u);
        const uint64_t size = bh >> 3;
        if (type == 3) return worst;
        total += 1;
        if (type == 1) {
            ip += 1;
        } else {
            if (n < ip + size) return worst;
            if (type == 2) {
                const uint8_t* p = in + ip;
                if (size < 1) return worst;
                const int ltype = p[0] & 3, sf = (p[0] >> 2) & 3;
                uint64_t hdr, comp;
                if (ltype < 2) {
                    uint64_t regen;
                    if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
                    else if (sf == 1) { if (size < 2) return worst; hdr = 2; regen = le(p, 2) >> 4; }
                    else { if (size < 3) return worst; hdr = 3; regen = le(p, 3) >> 4; }
                    comp = ltype == 0 ? regen : 1;
                } else {
                    if (sf == 0 || sf == 1) { if (size < 3) return worst; hdr = 3; comp = (le(p, 3) >> 14) & 0x3FF; }
                    else if (sf == 2) { if (size < 4) return worst; hdr = 4; comp = (le(p, 4) >> 18) & 0x3FFF; }
                    else { if (size < 5) return worst; hdr = 5; comp = (le(p, 5) >> 22) & 0x3FFFF; }
                }
                if (size < hdr + comp + 1) return worst;
                const uint8_t* q = p + hdr + comp;
                const uint64_t left = size - hdr - comp;
                uint64_t nseq;
                if (q[0] < 128) nseq = q[0];
                else if (q[0] < 255) { if (left < 2) return worst; nseq = ((uint64_t)(q[0] - 128) << 8) + q[1]; }
                else { if (left < 3) return worst; nseq = (uint64_t)q[1] + ((uint64_t)q[2] << 8) + 0x7F00; }
                total += nseq;
            }
            ip += size;
        }
        if (total >= worst) return worst;
        if (last) break;
    }
    return total;
}

// What the copy stage does, sequentially (host tests; the device runs l4_copy of lz4_block_cta.cuh over
// the same descriptors): out[0, total) from descriptors and literals.  Returns total or < 0.
FSB_HDN int64_t apply_descriptors(const SeqDesc* d, uint32_t nd, const uint8_t* lit, uint64_t lit_used, uint8_t* out,
                                  uint64_t total)
{
    for (uint32_t k = 0; k < nd; ++k) {
        const uint64_t o = d[k].out_pos, oe = k + 1u < nd ? d[k + 1u].out_pos : total;
        if (oe < o || oe > total || d[k].lit > oe - o || (uint64_t)d[k].lit_pos + d[k].lit > lit_used) return kErrOut;
        copy16(out + o, lit + d[k].lit_pos, d[k].lit);
        const uint64_t m = o + d[k].lit, ml = oe - m;
        if (ml) {
            if (d[k].off == 0 || d[k].off > m) return kErrSeq;
            copy_match(out + m, d[k].off, ml);
        }
    }
    return (int64_t)total;
}

}  // namespace zstd
}  // namespace fsb200
