/*
 * lz4_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).
 *
 * CPU restatement of what the reference's block-file loop does around the hot
 * path (benchmark/flagstats.cpp:288-358): walk a container of
 *     [int32 raw_size][int32 comp_size][LZ4 block]
 * records, decode every block (LZ4_decompress_safe, :316) and hand
 * N = raw_size >> 1 records (:323) to the flagstat kernel (:328-329).
 *
 * The decoder is the lz4 *library* the reference links as a system dependency
 * (`-llz4`, Makefile:32; no version pinned, headers not vendored, absent from
 * this image).  It is restated from the published "LZ4 Block Format
 * Description" (lz4_Block_format.md): token, literal-length / match-length
 * extension bytes (+255 while the byte is 255), little-endian 16-bit offset,
 * minimum match 4, overlapping matches repeat the last `offset` bytes, the last
 * sequence carries literals only.  Parity is pinned against a real liblz4
 * through pyarrow's "lz4_raw" codec in tests/test_lz4_oracle.py (both
 * directions: blocks compressed by liblz4 decode to the original here, and
 * blocks written by oracle_lz4_compress below decode with liblz4).
 */
#include <stdint.h>
#include <string.h>

/* returns bytes produced, or < 0 for a malformed block */
int64_t oracle_lz4_decompress(const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap)
{
    uint64_t ip = 0, op = 0;
    while (ip < in_size) {
        const unsigned token = in[ip++];
        uint64_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= in_size) return -1;
                b = in[ip++];
                lit += b;
            } while (b == 255);
        }
        if (lit > in_size - ip || lit > out_cap - op) return -2;
        memcpy(out + op, in + ip, lit);
        ip += lit;
        op += lit;
        if (ip >= in_size) break;
        if (in_size - ip < 2) return -3;
        const uint64_t offset = (uint64_t)in[ip] | ((uint64_t)in[ip + 1] << 8);
        ip += 2;
        uint64_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (ip >= in_size) return -1;
                b = in[ip++];
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (offset == 0 || offset > op || ml > out_cap - op) return -4;
        for (uint64_t i = 0; i < ml; ++i) out[op + i] = out[op - offset + i];  /* byte-serial: overlap repeats */
        op += ml;
    }
    return (int64_t)op;
}

/*
 * A plain greedy LZ4 block encoder (hash of 4 bytes, single candidate), only so
 * that tests can build containers whose blocks do not come from liblz4 and so
 * exercise long matches, overlapping matches and length-extension bytes on
 * purpose.  Produces valid blocks under the format's end-of-block rules (last 5
 * bytes are literals, last match starts >= 12 bytes before the end).
 * Returns the compressed size, or -1 if `cap` is too small.
 */
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

static int64_t emit_seq(uint8_t* out, uint64_t cap, uint64_t* po, const uint8_t* lit, uint64_t nlit,
                        uint64_t offset, uint64_t ml /* 0 = final literals */)
{
    uint64_t o = *po;
    uint64_t need = 1 + nlit / 255 + 1 + nlit + (ml ? 2 + (ml - 4) / 255 + 1 : 0);
    if (o + need > cap) return -1;
    const uint64_t mcode = ml ? ml - 4 : 0;
    out[o++] = (uint8_t)(((nlit >= 15 ? 15 : nlit) << 4) | (ml ? (mcode >= 15 ? 15 : mcode) : 0));
    if (nlit >= 15) { uint64_t r = nlit - 15; while (r >= 255) { out[o++] = 255; r -= 255; } out[o++] = (uint8_t)r; }
    memcpy(out + o, lit, nlit);
    o += nlit;
    if (ml) {
        out[o++] = (uint8_t)(offset & 255);
        out[o++] = (uint8_t)(offset >> 8);
        if (mcode >= 15) { uint64_t r = mcode - 15; while (r >= 255) { out[o++] = 255; r -= 255; } out[o++] = (uint8_t)r; }
    }
    *po = o;
    return 0;
}

int64_t oracle_lz4_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap)
{
    enum { HBITS = 16 };
    static uint32_t table[1 << HBITS];
    memset(table, 0xFF, sizeof(table));
    uint64_t o = 0, anchor = 0, i = 0;
    if (n >= 13) {
        const uint64_t mflimit = n - 12;  /* last match must start before this */
        while (i < mflimit) {
            const uint32_t h = (rd32(in + i) * 2654435761u) >> (32 - HBITS);
            const uint32_t cand = table[h];
            table[h] = (uint32_t)i;
            if (cand != 0xFFFFFFFFu && i - cand <= 65535 && rd32(in + cand) == rd32(in + i)) {
                uint64_t ml = 4;
                const uint64_t mend = n - 5;  /* last 5 bytes stay literals */
                while (i + ml < mend && in[cand + ml] == in[i + ml]) ++ml;
                if (emit_seq(out, cap, &o, in + anchor, i - anchor, i - cand, ml) < 0) return -1;
                i += ml;
                anchor = i;
            } else {
                ++i;
            }
        }
    }
    if (emit_seq(out, cap, &o, in + anchor, n - anchor, 0, 0) < 0) return -1;
    return (int64_t)o;
}

/*
 * The block loop of benchmark/flagstats.cpp:288-358 over a container held in
 * memory.  Decodes every block into `scratch` (>= the largest raw_size) and
 * calls `count(records, n, ctx)`.  Returns the number of records, or < 0.
 */
typedef void (*oracle_block_fn)(const uint16_t*, uint64_t, void*);

int64_t oracle_lz4_container_walk(const uint8_t* bytes, uint64_t n_bytes, uint8_t* scratch,
                                  uint64_t scratch_cap, oracle_block_fn count, void* ctx)
{
    uint64_t pos = 0, records = 0;
    while (pos < n_bytes) {
        int32_t raw, comp;
        if (n_bytes - pos < 8) return -10;
        memcpy(&raw, bytes + pos, 4);
        memcpy(&comp, bytes + pos + 4, 4);
        pos += 8;
        if (raw < 0 || comp <= 0 || (uint64_t)comp > n_bytes - pos || (uint64_t)raw > scratch_cap) return -11;
        const int64_t got = oracle_lz4_decompress(bytes + pos, (uint64_t)comp, scratch, (uint64_t)raw);
        if (got != raw) return -12;
        pos += (uint64_t)comp;
        if (count) count((const uint16_t*)scratch, (uint64_t)raw >> 1, ctx);  /* N = size >> 1, :323 */
        records += (uint64_t)raw >> 1;
    }
    return (int64_t)records;
}
