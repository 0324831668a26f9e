/*
 * zstd_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).
 *
 * CPU restatement of the decoder behind the reference's Zstd block files
 * (benchmark/flagstats.cpp:192-215 writes [int32 raw_size][int32 comp_size][Zstd frame]
 * records with ZSTD_compress, :90-93; zstd_decompress(), :636-676, decodes every record
 * with ZSTD_decompress, :95-98, and feeds N = raw_size >> 1 records to the flagstat kernel,
 * :661-665, exactly like the LZ4 reader).  The reference passes the size of its read buffer
 * (1,024,000) instead of comp_size as the source length (:659) and ignores the error that
 * produces once the first frame has been decoded; decoding exactly the one frame of
 * comp_size bytes, as done here, gives the same bytes.
 *
 * zstd is a system library the reference links (`-lzstd`, no version pinned, headers not
 * vendored and absent from this image; the runtime libzstd.so.1 IS present).  What is
 * restated here is the published format, RFC 8878 "Zstandard Compression and the
 * application/zstd Media Type": frame header, raw / RLE / compressed blocks, the literals
 * section (raw, RLE, Huffman with 1 or 4 streams, treeless), Huffman tree descriptions
 * (direct or FSE-compressed weights), the sequences section (predefined / RLE / FSE /
 * repeat tables, backward bitstream), repeat offsets and sequence execution.  No
 * dictionaries; the optional content checksum is skipped, not verified.  Parity is
 * pinned against the real libzstd (ctypes on libzstd.so.1, and pyarrow's codec) in
 * tests/test_zstd_oracle.py: frames written by ZSTD_compress at every level the
 * reference's table uses (README.md:148-175) must decode here to the original bytes.
 *
 * Written from the RFC, not from zstd's sources.  Error returns are negative; nothing is
 * read or written out of bounds for any input (tests/test_oracle_sanitizers.py fuzzes it).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZE_TRUNC (-1)    /* input ends early */
#define ZE_MAGIC (-2)
#define ZE_HEADER (-3)   /* reserved bits, dictionary, window */
#define ZE_BLOCK (-4)
#define ZE_LITERALS (-5)
#define ZE_HUF (-6)
#define ZE_FSE (-7)
#define ZE_SEQ (-8)
#define ZE_OUT (-9)      /* output capacity / declared size mismatch */

/* ---- bit readers ------------------------------------------------------------ */

/* forward, little-endian bit order (FSE table descriptions) */
typedef struct { const uint8_t* p; uint64_t nbits; uint64_t pos; } FwdBits;

static uint32_t fwd_read(FwdBits* b, int n)
{
    uint32_t v = 0;
    for (int i = 0; i < n; ++i, ++b->pos)
        if (b->pos < b->nbits) v |= (uint32_t)((b->p[b->pos >> 3] >> (b->pos & 7)) & 1u) << i;
    return v;
}

/* backward bitstream: the last byte carries a 1-bit end marker above the payload */
typedef struct { const uint8_t* p; int64_t pos; } BackBits;

static int back_init(BackBits* b, const uint8_t* p, uint64_t n)
{
    if (n == 0 || p[n - 1] == 0) return -1;
    int hb = 7;
    while (!((p[n - 1] >> hb) & 1)) --hb;
    b->p = p;
    b->pos = (int64_t)(n - 1) * 8 + hb;
    return 0;
}

/* n <= 32 bits ending at the current position; bits before the start read as 0 */
static uint64_t back_read(BackBits* b, int n)
{
    uint64_t v = 0;
    b->pos -= n;
    for (int i = 0; i < n; ++i) {
        const int64_t q = b->pos + i;
        if (q >= 0) v |= (uint64_t)((b->p[q >> 3] >> (q & 7)) & 1u) << i;
    }
    return v;
}

static int highbit(uint32_t v)  /* floor(log2(v)), v > 0 */
{
    int h = 0;
    while (v >>= 1) ++h;
    return h;
}

/* ---- FSE ---------------------------------------------------------------------- */

#define FSE_MAX_LOG 9
typedef struct {
    int log;                             /* accuracy log; -1 = no table yet */
    uint8_t sym[1 << FSE_MAX_LOG];
    uint8_t nbits[1 << FSE_MAX_LOG];
    uint16_t base[1 << FSE_MAX_LOG];
} FseTable;

/* RFC 8878 4.1.1: normalised counts -> decoding table */
static int fse_build(FseTable* t, const int16_t* freq, int nsym, int log)
{
    const int size = 1 << log;
    uint16_t next[256];
    int high = size;
    if (log > FSE_MAX_LOG || nsym > 256) return ZE_FSE;
    for (int s = 0; s < nsym; ++s)
        if (freq[s] == -1) {
            t->sym[--high] = (uint8_t)s;
            next[s] = 1;
        }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; ++s) {
        if (freq[s] <= 0) continue;
        next[s] = (uint16_t)freq[s];
        for (int i = 0; i < freq[s]; ++i) {
            t->sym[pos] = (uint8_t)s;
            do pos = (pos + step) & mask; while (pos >= high);
        }
    }
    if (pos != 0) return ZE_FSE;
    for (int i = 0; i < size; ++i) {
        const uint16_t n = next[t->sym[i]]++;
        t->nbits[i] = (uint8_t)(log - highbit(n));
        t->base[i] = (uint16_t)(((uint32_t)n << t->nbits[i]) - (uint32_t)size);
    }
    t->log = log;
    return 0;
}

/* RFC 8878 4.1.1: FSE table description; returns bytes consumed or < 0 */
static int64_t fse_read_table(FseTable* t, const uint8_t* p, uint64_t n, int max_log, int max_sym)
{
    FwdBits b = {p, n * 8, 0};
    int16_t freq[256];
    const int log = 5 + (int)fwd_read(&b, 4);
    if (log > max_log) return ZE_FSE;
    int remaining = 1 << log, s = 0;
    while (remaining > 0 && s <= max_sym) {
        const int bits = highbit((uint32_t)remaining + 1) + 1;
        uint32_t v = fwd_read(&b, bits);
        const uint32_t lower = (1u << (bits - 1)) - 1u;
        const uint32_t thresh = (1u << bits) - 1u - ((uint32_t)remaining + 1u);
        if ((v & lower) < thresh) {
            b.pos -= 1;  /* the value took one bit less */
            v &= lower;
        } else if (v > lower) {
            v -= thresh;
        }
        const int proba = (int)v - 1;
        remaining -= proba < 0 ? 1 : proba;
        freq[s++] = (int16_t)proba;
        if (proba == 0) {
            for (;;) {
                const int rep = (int)fwd_read(&b, 2);
                for (int i = 0; i < rep && s <= max_sym; ++i) freq[s++] = 0;
                if (rep != 3) break;
            }
        }
        if (b.pos > b.nbits) return ZE_TRUNC;
    }
    if (remaining != 0 || s > max_sym + 1) return ZE_FSE;
    const int rc = fse_build(t, freq, s, log);
    if (rc) return rc;
    return (int64_t)((b.pos + 7) >> 3);
}

static void fse_rle(FseTable* t, uint8_t sym)
{
    t->log = 0;
    t->sym[0] = sym;
    t->nbits[0] = 0;
    t->base[0] = 0;
}

/* ---- Huffman (RFC 8878 4.2) ------------------------------------------------------ */

#define HUF_MAX_BITS 11
typedef struct {
    int bits;                            /* table log; 0 = no table yet */
    uint8_t sym[1 << HUF_MAX_BITS];
    uint8_t len[1 << HUF_MAX_BITS];
} HufTable;

static int huf_build(HufTable* h, uint8_t* w, int n)  /* w[0..n-1] given, w[n] implied */
{
    uint32_t sum = 0;
    for (int i = 0; i < n; ++i) {
        if (w[i] > HUF_MAX_BITS) return ZE_HUF;
        if (w[i]) sum += 1u << (w[i] - 1);
    }
    if (sum == 0) return ZE_HUF;
    const int max_bits = highbit(sum) + 1;
    const uint32_t left = (1u << max_bits) - sum;
    if (max_bits > HUF_MAX_BITS || (left & (left - 1)) != 0 || n >= 256) return ZE_HUF;
    w[n] = (uint8_t)(highbit(left) + 1);
    ++n;
    /* canonical assignment: by increasing weight, then by symbol value */
    uint32_t rank_start[HUF_MAX_BITS + 2] = {0};
    for (int i = 0; i < n; ++i)
        if (w[i]) rank_start[w[i] + 1] += 1u << (w[i] - 1);
    for (int k = 1; k <= HUF_MAX_BITS + 1; ++k) rank_start[k] += rank_start[k - 1];
    for (int i = 0; i < n; ++i) {
        if (!w[i]) continue;
        const uint32_t span = 1u << (w[i] - 1);
        const uint8_t len = (uint8_t)(max_bits + 1 - w[i]);
        for (uint32_t k = 0; k < span; ++k) {
            h->sym[rank_start[w[i]] + k] = (uint8_t)i;
            h->len[rank_start[w[i]] + k] = len;
        }
        rank_start[w[i]] += span;
    }
    h->bits = max_bits;
    return 0;
}

/* tree description; returns bytes consumed or < 0 */
static int64_t huf_read_tree(HufTable* h, const uint8_t* p, uint64_t n)
{
    uint8_t w[257];
    if (n < 1) return ZE_TRUNC;
    const int hb = p[0];
    if (hb >= 128) {  /* direct: 4 bits per weight */
        const int cnt = hb - 127;
        const uint64_t bytes = ((uint64_t)cnt + 1) / 2;
        if (n < 1 + bytes) return ZE_TRUNC;
        for (int i = 0; i < cnt; ++i) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        const int rc = huf_build(h, w, cnt);
        return rc ? rc : (int64_t)(1 + bytes);
    }
    if (hb == 0 || n < 1 + (uint64_t)hb) return ZE_TRUNC;
    FseTable t;
    const int64_t used = fse_read_table(&t, p + 1, (uint64_t)hb, 6, 255);
    if (used < 0) return used;
    if (used >= hb) return ZE_HUF;
    BackBits b;
    if (back_init(&b, p + 1 + used, (uint64_t)hb - (uint64_t)used)) return ZE_HUF;
    uint32_t s1 = (uint32_t)back_read(&b, t.log), s2 = (uint32_t)back_read(&b, t.log);
    if (b.pos < 0) return ZE_HUF;
    int cnt = 0;
    for (;;) {  /* two interleaved states until the stream runs dry */
        if (cnt >= 254) return ZE_HUF;
        w[cnt++] = t.sym[s1];
        s1 = t.base[s1] + (uint32_t)back_read(&b, t.nbits[s1]);
        if (b.pos < 0) {
            w[cnt++] = t.sym[s2];
            break;
        }
        if (cnt >= 254) return ZE_HUF;
        w[cnt++] = t.sym[s2];
        s2 = t.base[s2] + (uint32_t)back_read(&b, t.nbits[s2]);
        if (b.pos < 0) {
            w[cnt++] = t.sym[s1];
            break;
        }
    }
    const int rc = huf_build(h, w, cnt);
    return rc ? rc : (int64_t)(1 + hb);
}

/* one Huffman stream -> exactly `want` symbols */
static int huf_stream(const HufTable* h, const uint8_t* p, uint64_t n, uint8_t* out, uint64_t want)
{
    BackBits b;
    if (back_init(&b, p, n)) return ZE_HUF;
    const uint32_t mask = (1u << h->bits) - 1u;
    uint32_t state = (uint32_t)back_read(&b, h->bits);
    uint64_t got = 0;
    while (b.pos > -(int64_t)h->bits) {
        if (got == want) return ZE_HUF;
        out[got++] = h->sym[state];
        const int l = h->len[state];
        state = ((state << l) + (uint32_t)back_read(&b, l)) & mask;
    }
    if (b.pos != -(int64_t)h->bits || got != want) return ZE_HUF;
    return 0;
}

/* ---- sequences (RFC 8878 3.1.1.3.2) -------------------------------------------------- */

static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28,
                                     32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2,
                                    3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24,
                                     25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83,
                                     99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4,
                                    5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2,
                                       2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, -1, -1, -1, -1, -1};

typedef struct {
    FseTable ll, of, ml;
    HufTable huf;
    uint64_t rep[3];
} FrameCtx;

/* one of the three tables of a sequences section; returns bytes consumed or < 0 */
static int64_t seq_table(FseTable* t, int mode, const uint8_t* p, uint64_t n, const int16_t* def, int def_n,
                         int def_log, int max_log, int max_sym)
{
    switch (mode) {
        case 0: {
            const int rc = fse_build(t, def, def_n, def_log);
            return rc ? rc : 0;
        }
        case 1:
            if (n < 1) return ZE_TRUNC;
            if (p[0] > max_sym) return ZE_SEQ;
            fse_rle(t, p[0]);
            return 1;
        case 2:
            return fse_read_table(t, p, n, max_log, max_sym);
        default:
            return t->log < 0 ? ZE_SEQ : 0;  /* repeat: the previous table must exist */
    }
}

/* a compressed block: literals + sequences -> out[op..]; returns new op or < 0 */
static int64_t block_compressed(FrameCtx* c, const uint8_t* p, uint64_t n, uint8_t* out, uint64_t op, uint64_t cap,
                                uint8_t* lit_buf /* >= 128 KiB + 32 */)
{
    /* ---- literals section ---- */
    if (n < 1) return ZE_TRUNC;
    const int ltype = p[0] & 3, sf = (p[0] >> 2) & 3;
    uint64_t regen, comp = 0, hdr;
    int streams = 1;
    const uint8_t* lit = NULL;
    if (ltype < 2) {  /* raw / RLE */
        if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return ZE_TRUNC; hdr = 2; regen = (p[0] >> 4) | ((uint64_t)p[1] << 4); }
        else { if (n < 3) return ZE_TRUNC; hdr = 3; regen = (p[0] >> 4) | ((uint64_t)p[1] << 4) | ((uint64_t)p[2] << 12); }
        if (regen > (128u << 10)) return ZE_LITERALS;
        if (ltype == 0) {
            if (n < hdr + regen) return ZE_TRUNC;
            lit = p + hdr;
            comp = regen;
        } else {
            if (n < hdr + 1) return ZE_TRUNC;
            memset(lit_buf, p[hdr], regen);
            lit = lit_buf;
            comp = 1;
        }
    } else {  /* Huffman-compressed / treeless */
        if (sf == 0 || sf == 1) {
            if (n < 3) return ZE_TRUNC;
            hdr = 3;
            const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
            regen = (v >> 4) & 0x3FF;
            comp = (v >> 14) & 0x3FF;
            streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return ZE_TRUNC;
            hdr = 4;
            const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            regen = (v >> 4) & 0x3FFF;
            comp = (v >> 18) & 0x3FFF;
            streams = 4;
        } else {
            if (n < 5) return ZE_TRUNC;
            hdr = 5;
            const uint64_t v = (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24) |
                               ((uint64_t)p[4] << 32);
            regen = (v >> 4) & 0x3FFFF;
            comp = (v >> 22) & 0x3FFFF;
            streams = 4;
        }
        if (regen > (128u << 10) || n < hdr + comp) return ZE_LITERALS;
        const uint8_t* q = p + hdr;
        uint64_t left = comp;
        if (ltype == 2) {
            const int64_t used = huf_read_tree(&c->huf, q, left);
            if (used < 0) return used;
            q += used;
            left -= (uint64_t)used;
        } else if (c->huf.bits == 0) {
            return ZE_HUF;  /* treeless without a previous tree */
        }
        if (streams == 1) {
            const int rc = huf_stream(&c->huf, q, left, lit_buf, regen);
            if (rc) return rc;
        } else {
            if (left < 6) return ZE_LITERALS;
            const uint64_t s1 = q[0] | ((uint64_t)q[1] << 8), s2 = q[2] | ((uint64_t)q[3] << 8),
                           s3 = q[4] | ((uint64_t)q[5] << 8);
            if (6 + s1 + s2 + s3 > left) return ZE_LITERALS;
            const uint64_t s4 = left - 6 - s1 - s2 - s3;
            const uint64_t each = (regen + 3) / 4;
            if (3 * each > regen) return ZE_LITERALS;
            int rc;
            if ((rc = huf_stream(&c->huf, q + 6, s1, lit_buf, each))) return rc;
            if ((rc = huf_stream(&c->huf, q + 6 + s1, s2, lit_buf + each, each))) return rc;
            if ((rc = huf_stream(&c->huf, q + 6 + s1 + s2, s3, lit_buf + 2 * each, each))) return rc;
            if ((rc = huf_stream(&c->huf, q + 6 + s1 + s2 + s3, s4, lit_buf + 3 * each, regen - 3 * each))) return rc;
        }
        lit = lit_buf;
    }
    p += hdr + comp;
    n -= hdr + comp;

    /* ---- sequences section ---- */
    if (n < 1) return ZE_TRUNC;
    uint64_t nseq;
    if (p[0] == 0) { nseq = 0; p += 1; n -= 1; }
    else if (p[0] < 128) { nseq = p[0]; p += 1; n -= 1; }
    else if (p[0] < 255) { if (n < 2) return ZE_TRUNC; nseq = ((uint64_t)(p[0] - 128) << 8) + p[1]; p += 2; n -= 2; }
    else { if (n < 3) return ZE_TRUNC; nseq = (uint64_t)p[1] + ((uint64_t)p[2] << 8) + 0x7F00; p += 3; n -= 3; }
    uint64_t lp = 0;  /* literals consumed */
    if (nseq) {
        if (n < 1) return ZE_TRUNC;
        const int modes = p[0];
        if (modes & 3) return ZE_SEQ;
        p += 1; n -= 1;
        int64_t used;
        if ((used = seq_table(&c->ll, modes >> 6, p, n, LL_DEFAULT, 36, 6, 9, 35)) < 0) return used;
        p += used; n -= (uint64_t)used;
        if ((used = seq_table(&c->of, (modes >> 4) & 3, p, n, OF_DEFAULT, 29, 5, 8, 31)) < 0) return used;
        p += used; n -= (uint64_t)used;
        if ((used = seq_table(&c->ml, (modes >> 2) & 3, p, n, ML_DEFAULT, 53, 6, 9, 52)) < 0) return used;
        p += used; n -= (uint64_t)used;
        BackBits b;
        if (back_init(&b, p, n)) return ZE_SEQ;
        uint32_t sl = (uint32_t)back_read(&b, c->ll.log), so = (uint32_t)back_read(&b, c->of.log),
                 sm = (uint32_t)back_read(&b, c->ml.log);
        for (uint64_t i = 0; i < nseq; ++i) {
            const int oc = c->of.sym[so], mc = c->ml.sym[sm], lc = c->ll.sym[sl];
            if (oc > 31 || mc > 52 || lc > 35) return ZE_SEQ;
            const uint64_t ov = ((uint64_t)1 << oc) + back_read(&b, oc);
            const uint64_t ml = ML_BASE[mc] + back_read(&b, ML_BITS[mc]);
            const uint64_t ll = LL_BASE[lc] + back_read(&b, LL_BITS[lc]);
            if (i + 1 < nseq) {
                sl = c->ll.base[sl] + (uint32_t)back_read(&b, c->ll.nbits[sl]);
                sm = c->ml.base[sm] + (uint32_t)back_read(&b, c->ml.nbits[sm]);
                so = c->of.base[so] + (uint32_t)back_read(&b, c->of.nbits[so]);
            }
            if (b.pos < 0) return ZE_SEQ;
            /* repeat offsets, RFC 8878 3.1.1.5 */
            uint64_t off;
            if (ov > 3) {
                off = ov - 3;
                c->rep[2] = c->rep[1]; c->rep[1] = c->rep[0]; c->rep[0] = off;
            } else {
                uint64_t idx = ov - 1 + (ll == 0 ? 1 : 0);
                if (idx == 0) {
                    off = c->rep[0];
                } else {
                    off = idx < 3 ? c->rep[idx] : c->rep[0] - 1;
                    if (idx > 1) c->rep[2] = c->rep[1];
                    c->rep[1] = c->rep[0];
                    c->rep[0] = off;
                }
            }
            /* execute: literals, then the match (may overlap its own output) */
            if (ll > regen - lp || ll + ml > cap - op) return ZE_OUT;
            memcpy(out + op, lit + lp, ll);
            op += ll;
            lp += ll;
            if (off == 0 || off > op) return ZE_SEQ;
            for (uint64_t k = 0; k < ml; ++k) out[op + k] = out[op + k - off];
            op += ml;
        }
        if (b.pos != 0) return ZE_SEQ;
    }
    if (regen - lp > cap - op) return ZE_OUT;
    memcpy(out + op, lit + lp, regen - lp);
    return (int64_t)(op + (regen - lp));
}

/* One frame (what ZSTD_compress writes, what ZSTD_decompress reads: benchmark/flagstats.cpp:90-98).
 * Returns bytes produced or < 0. */
int64_t oracle_zstd_decompress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap)
{
    if (n < 6) return ZE_TRUNC;
    uint32_t magic;
    memcpy(&magic, in, 4);
    if (magic != 0xFD2FB528u) return ZE_MAGIC;
    const int fhd = in[4];
    const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, dict = fhd & 3;
    if (fhd & 0x08) return ZE_HEADER;  /* reserved bit */
    uint64_t ip = 5;
    if (!single) ip += 1;  /* window descriptor: the whole frame is decoded into one buffer anyway */
    static const int dict_bytes[4] = {0, 1, 2, 4};
    if (dict) {
        if (n < ip + (uint64_t)dict_bytes[dict]) return ZE_TRUNC;
        uint32_t id = 0;
        memcpy(&id, in + ip, (size_t)dict_bytes[dict]);
        if (id != 0) return ZE_HEADER;  /* dictionaries are not supported */
        ip += (uint64_t)dict_bytes[dict];
    }
    const int fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (n < ip + (uint64_t)fcs_bytes) return ZE_TRUNC;
    uint64_t fcs = 0;
    memcpy(&fcs, in + ip, (size_t)fcs_bytes);
    if (fcs_bytes == 2) fcs += 256;
    ip += (uint64_t)fcs_bytes;
    if (fcs_bytes && fcs > cap) return ZE_OUT;

    FrameCtx* c = (FrameCtx*)malloc(sizeof(FrameCtx));
    uint8_t* lit_buf = (uint8_t*)malloc((128u << 10) + 32);
    if (!c || !lit_buf) { free(c); free(lit_buf); return ZE_OUT; }
    c->ll.log = c->of.log = c->ml.log = -1;
    c->huf.bits = 0;
    c->rep[0] = 1; c->rep[1] = 4; c->rep[2] = 8;
    int64_t rc = 0;
    uint64_t op = 0;
    for (;;) {
        if (n < ip + 3) { rc = ZE_TRUNC; break; }
        const uint32_t bh = (uint32_t)in[ip] | ((uint32_t)in[ip + 1] << 8) | ((uint32_t)in[ip + 2] << 16);
        ip += 3;
        const int last = (int)(bh & 1u), type = (int)((bh >> 1) & 3u);
        const uint64_t size = bh >> 3;
        if (type == 0) {
            if (n < ip + size) { rc = ZE_TRUNC; break; }
            if (size > cap - op) { rc = ZE_OUT; break; }
            memcpy(out + op, in + ip, size);
            op += size;
            ip += size;
        } else if (type == 1) {
            if (n < ip + 1) { rc = ZE_TRUNC; break; }
            if (size > cap - op) { rc = ZE_OUT; break; }
            memset(out + op, in[ip], size);
            op += size;
            ip += 1;
        } else if (type == 2) {
            if (size > (128u << 10) || n < ip + size) { rc = size > (128u << 10) ? ZE_BLOCK : ZE_TRUNC; break; }
            const int64_t r = block_compressed(c, in + ip, size, out, op, cap, lit_buf);
            if (r < 0) { rc = r; break; }
            op = (uint64_t)r;
            ip += size;
        } else {
            rc = ZE_BLOCK;
            break;
        }
        if (last) break;
    }
    free(c);
    free(lit_buf);
    if (rc < 0) return rc;
    if (checksum) {
        if (n < ip + 4) return ZE_TRUNC;
        ip += 4;  /* xxh64 of the content, low 32 bits: not verified here */
    }
    if (fcs_bytes && fcs != op) return ZE_OUT;
    return (int64_t)op;
}

/*
 * The block loop of the reference's zstd reader over a container held in memory:
 * [int32 raw_size][int32 comp_size][Zstd frame] ... (benchmark/flagstats.cpp:192-215 writes
 * them, :636-676 reads them); every frame is decoded into `scratch` and N = raw_size >> 1 records go to `count`.
 */
typedef void (*oracle_zblock_fn)(const uint16_t*, uint64_t, void*);

int64_t oracle_zstd_container_walk(const uint8_t* bytes, uint64_t n_bytes, uint8_t* scratch, uint64_t scratch_cap,
                                   oracle_zblock_fn count, void* ctx)
{
    uint64_t pos = 0, records = 0;
    while (pos < n_bytes) {
        int32_t raw, comp;
        if (n_bytes - pos < 8) return -10;
        memcpy(&raw, bytes + pos, 4);
        memcpy(&comp, bytes + pos + 4, 4);
        pos += 8;
        if (raw < 0 || comp <= 0 || (uint64_t)comp > n_bytes - pos || (uint64_t)raw > scratch_cap) return -11;
        const int64_t got = oracle_zstd_decompress(bytes + pos, (uint64_t)comp, scratch, (uint64_t)raw);
        if (got != raw) return -12;
        pos += (uint64_t)comp;
        if (count) count((const uint16_t*)scratch, (uint64_t)raw >> 1, ctx);
        records += (uint64_t)raw >> 1;
    }
    return (int64_t)records;
}
