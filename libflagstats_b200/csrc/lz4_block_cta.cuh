// lz4_block_cta.cuh -- LZ4 block decoder, ONE CTA (512 threads) PER BLOCK, two phases.
//
// The two warp-per-block decoders (lz4_block.cuh, lz4_block_group.cuh) take 25-30 ms for ONE
// 1,024,000-byte block of FLAG words whatever the file size: a block is one serial chain for one
// warp, and a phase profile of the group decoder (tools/lz4_phase_probe.py,
// profiles/r4j_lz4_phases.jsonl) puts 60-80 % of its cycles into the match machinery of 32
// sequences at a time.  A file of 400 blocks keeps 400 warps busy (2 % of the machine) and loses
// to the reference's loop on 16 host cores.  This decoder puts a whole CTA on a block:
//
//   PHASE A  parse -- no output byte is touched; the token chain only depends on the input.
//     The input is cut into 256-byte WINDOWS, sixteen per super-step, one per warp.  For every byte
//     position p of its window a warp computes where a sequence whose token sits at p would end
//     (nx0[p]; all 256 candidates at once, 8 per lane), then by pointer doubling over the
//     in-window successors (7 levels) the LAST candidate start a chain entering at p reaches
//     inside the window, hence exit[p] = where that chain leaves the window.  With the eight
//     exit maps in shared memory one thread chains the super-step's true entry points
//     (16 dependent look-ups instead of ~1400 dependent token parses), every warp then
//     enumerates the real sequence starts of its window from its entry (lane k = k-th successor
//     by binary decomposition of the same tables), parses them, and a scan over the output
//     lengths turns them into DESCRIPTORS {output position, token position}, 8 bytes each,
//     in a global scratch array (L2-resident).  Sequences whose lengths use more than four
//     extension bytes leave the window scheme ("escape") and are parsed by one thread.
//   PHASE B  copy -- the output is produced in TILES of 8 KiB by all 512 threads.  The tile's
//     descriptors and the input bytes they cover are staged in shared memory.  B1: one thread
//     per SEQUENCE stores its literal bytes and gives every match byte a PARENT -- the output
//     byte it copies (two parents per 32-bit store); parents before the tile are copied at once
//     from the 56 KiB of history the CTA keeps in shared memory (an LZ4 offset is < 65536; the
//     rare longer reach reads global memory); matches longer than 64 bytes go on a list that the
//     warps then work off together, 32 bytes at a time, so that no lane waits behind a long
//     match.  (The first version gave every thread 16 consecutive output bytes and let it walk
//     the sequences under them: 5.6 warp instructions per output byte, 42 % of the kernel;
//     profiles/r4q_lz4_phases_cta.jsonl.)  B3: every thread takes 16 consecutive bytes' parents
//     into registers and pointer jumping over the whole tile resolves chains of matches in
//     log3(depth) rounds; B4 copies root -> byte.  The tile leaves shared memory as coalesced
//     16-byte stores.
//
// Same block format, same Lz4BlockDesc interface, negative status for anything malformed; every
// position derived from the input is checked before it is used as an index.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lz4_block.cuh"

namespace fsb200 {

constexpr int kL4Threads = 512;
constexpr int kL4Warps = kL4Threads / 32;
constexpr uint32_t kL4Win = 256;              // input bytes per window
constexpr uint32_t kL4Stage = kL4Win + 32;    // a window sees 32 bytes beyond its end: a simple sequence is <= 18 + 4 bytes long
constexpr uint32_t kL4Levels = 7;             // 2^7 = 128 > 85 = most sequences (>= 3 bytes) in a window
constexpr uint32_t kL4Tile = 8192;            // output bytes per tile
constexpr uint32_t kL4Ring = 65536;           // the tile being built + 56 KiB of history (a power of two: index = v & mask;
                                              // the rare match that reaches further back reads global memory)
constexpr uint32_t kL4MaxExt = 4;             // extension bytes per length the window scheme follows
constexpr uint32_t kL4ByteChunk = kL4Tile / kL4Threads;  // 16 output bytes per thread and tile
constexpr uint32_t kL4LongMatch = 64;                    // match bytes (inside the tile) above which the warps share a match
constexpr uint32_t kL4MaxLong = kL4Tile / kL4LongMatch;  // ... of which a tile holds at most this many

// nx0 codes (u16): < 0xFFF0 = successor position relative to the window start
constexpr uint32_t kL4Last = 0xFFF1;  // the token's literals end exactly at the end of the input: last sequence
constexpr uint32_t kL4Esc = 0xFFF2;   // needs the serial parser (long length fields)
constexpr uint32_t kL4Bad = 0xFFF3;   // runs past the end of the input

// Phase profile (tools/lz4_phase_probe.py, -DFSB_LZ4_PROFILE; never in the product): thread 0 of the
// CTA sums clock64() deltas per phase into g_lz4_prof[16 + k].
#ifdef FSB_LZ4_PROFILE
__device__ unsigned long long g_l4_prof[16];
#define L4P_DECL long long l4p_t = clock64()
#define L4P(k)                                                                     \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            const long long now_ = clock64();                                      \
            atomicAdd(&g_l4_prof[k], (unsigned long long)(now_ - l4p_t));          \
            l4p_t = now_;                                                          \
        }                                                                          \
    } while (0)
#define L4P_COUNT(k, v)                                                            \
    do {                                                                           \
        if (threadIdx.x == 0) atomicAdd(&g_l4_prof[k], (unsigned long long)(v));   \
    } while (0)
#define L4P_RESET l4p_t = clock64()
#else
#define L4P_DECL do { } while (0)
#define L4P(k) do { } while (0)
#define L4P_COUNT(k, v) do { } while (0)
#define L4P_RESET do { } while (0)
#endif

struct __align__(16) L4Desc {
    uint32_t out_pos;  // first output byte of the sequence (its literals)
    uint32_t lit_pos;  // position of its first literal byte in the input
    uint32_t lit;      // literal bytes; the match follows them and ends where the next sequence starts
    uint32_t off;      // match offset (0: the block's last sequence, literals only)
};

// shared memory: | ring 64 KiB (phase A: per-warp parse tables + the staged super-step) | P 16 KiB |
// scalars | long matches 2 KiB | descriptors of one tile 28 KiB | = 112,896 bytes: two CTAs per SM.  The
// descriptors of tile t + 1 travel (cp.async) while tile t is being resolved; literal bytes -- rare in FLAG
// data -- are read straight from global memory.
constexpr uint32_t kL4ParsePerWarp = 2u * kL4Win /*nx0 u16*/ + kL4Levels * kL4Win /*f[L] u8*/ + 2u * kL4Win /*exit u16*/;  // 2816
constexpr uint32_t kL4SuperStage = kL4Warps * kL4Win + 32u + 16u;  // the super-step's input bytes (+ alignment slack)
static_assert(kL4ParsePerWarp * kL4Warps + kL4SuperStage + 16u <= kL4Ring, "parse tables must fit into the ring area");
struct L4Long {  // a long match (or run of literals), clipped to the tile (v-space)
    uint32_t a, b;  // its bytes inside the tile: [a, b)
    uint32_t m;     // first byte of the match; literals: input position of byte v is v + m
    uint32_t off;   // 0: literals
};
constexpr uint32_t kL4DescStage = 1792;  // descriptors of one tile staged in shared memory (a tile of FLAG data has 500 - 1300)
constexpr size_t kL4Smem =
    (size_t)kL4Ring + 2u * kL4Tile + 256u + sizeof(L4Long) * kL4MaxLong + sizeof(L4Desc) * kL4DescStage;

// scratch one CTA needs for blocks of at most max_comp compressed / max_raw decoded bytes
__host__ __device__ inline size_t l4_scratch_bytes(uint32_t max_comp, uint32_t max_raw)
{
    const size_t descs = (size_t)max_comp / 3u + 64u;                  // a sequence with a match is >= 3 bytes
    const size_t tiles = ((size_t)max_raw + 15u) / kL4Tile + 4u;
    return ((descs * sizeof(L4Desc) + tiles * 4u) + 255u) & ~(size_t)255u;
}

struct L4Shared {  // the scalars at the end of the dynamic shared memory
    int err;
    uint32_t entry[kL4Warps];   // entry offset of window w in this super-step, 0xFFFFFFFF = skipped
    uint32_t cnt[kL4Warps];     // sequences of window w
    uint32_t outlen[kL4Warps];  // output bytes of window w
    uint32_t stop_code;         // kL4Last / kL4Esc / kL4Bad / 0
    uint32_t stop_pos;          // absolute token position of the sequence that stopped the chain
    uint32_t next_pos;          // where the next super-step starts
    uint32_t pos, out, nseq;    // running state of phase A
    uint32_t done;
    uint32_t n_long;            // phase B: long matches of this tile
};
static_assert(sizeof(L4Shared) <= 256, "the scalars have 256 bytes");

// Length fields of the sequence whose token is at `p` (absolute), read through `rd(i)` = input
// byte i (caller guarantees i < in_size).  Follows at most `max_ext` extension bytes per field.
// Returns false if a field needs more (escape) -- only possible when max_ext is finite.
// lit_pos = first literal byte.  Does NOT read the match fields (see l4_match_len).
template <class Rd>
__device__ __forceinline__ int l4_literal_len(Rd rd, uint32_t p, uint32_t in_size, uint32_t max_ext, uint32_t& lit,
                                              uint32_t& lit_pos)
{
    const uint32_t tok = rd(p);
    lit = tok >> 4;
    uint32_t q = p + 1u;
    if (lit == 15u) {
        uint32_t n = 0u, b;
        do {
            if (q >= in_size) return (int)kL4Bad;
            if (n == max_ext) return (int)kL4Esc;
            b = rd(q++);
            lit += b;
            ++n;
        } while (b == 255u);
    }
    lit_pos = q;
    return 0;
}

// Phase A for one block.  Writes descriptors and the per-tile index; returns the number of
// sequences, or a negative error.  *total_out receives the decoded size.
__device__ int l4_parse(const uint8_t* __restrict__ in, uint32_t in_size, uint32_t out_cap, L4Desc* __restrict__ desc,
                        uint32_t desc_cap, unsigned char* parse_smem, L4Shared* sh, uint32_t* total_out)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned char* mine = parse_smem + warp * kL4ParsePerWarp;
    uint16_t* nx0 = reinterpret_cast<uint16_t*>(mine);                       // [256]
    uint8_t* f = mine + 2u * kL4Win;                                         // [levels][256]
    uint16_t* exitp = reinterpret_cast<uint16_t*>(f + kL4Levels * kL4Win);   // [256]
    uint8_t* sst = parse_smem + kL4Warps * kL4ParsePerWarp;                  // the super-step's input, 16-byte aligned
    constexpr uint32_t kFull = 0xffffffffu;

    if (tid == 0) {
        sh->pos = 0u;
        sh->out = 0u;
        sh->nseq = 0u;
        sh->done = 0u;
    }
    __syncthreads();

    L4P_DECL;
    for (;;) {
        const uint32_t B = sh->pos;  // start of this super-step: a true sequence start
        if (sh->err) return sh->err;
        if (sh->done) break;
        L4P_COUNT(12, 1);  // super-steps
        // ---- 1. per warp: successor of every position of its window, exit map ----------------------
        // the super-step's input [B, B + 16 * 256 + 32) with aligned 16-byte loads (al = B's address mod 16;
        // bytes in front of the block belong to the same buffer), zero behind the end of the input
        const uint32_t al = (uint32_t)((reinterpret_cast<uintptr_t>(in) + B) & 15u);
        {
            const uint32_t nvec = (al + kL4Warps * kL4Win + 32u + 15u) >> 4;
            for (uint32_t v = tid; v < nvec; v += kL4Threads) {
                const long long rel = (long long)B - (long long)al + (long long)(v << 4);  // first byte, relative to `in`
                if (rel + 16 <= (long long)in_size) {
                    *reinterpret_cast<uint4*>(sst + (v << 4)) = *reinterpret_cast<const uint4*>(in + rel);
                } else {
                    for (uint32_t i = 0; i < 16u; ++i)
                        sst[(v << 4) + i] = rel + (long long)i < (long long)in_size ? in[rel + i] : (uint8_t)0;
                }
            }
        }
        __syncthreads();
        const uint32_t wb = B + warp * kL4Win;
        const bool have = wb < in_size;
        const uint8_t* stage = sst + al + warp * kL4Win;  // this warp's window and the 32 bytes behind it
        if (have) {
            const uint32_t avail = in_size - wb;
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) {
                const uint32_t p = lane + 32u * j;
                uint32_t code;
                const uint32_t tok0 = stage[p];
                if (p >= avail) {
                    code = kL4Bad;
                } else if ((tok0 >> 4) != 15u && (tok0 & 15u) != 15u) {
                    // the common token: no extension bytes, the successor sits right behind the offset
                    const uint32_t e = wb + p + 1u + (tok0 >> 4);  // end of the literals
                    code = e + 2u <= in_size ? p + 3u + (tok0 >> 4) : e == in_size ? kL4Last : kL4Bad;
                } else {
                    // a candidate whose fields reach past the staged bytes is an escape: the serial
                    // parser reads it from global memory if the chain really gets there
                    auto rd = [&](uint32_t i) -> uint32_t { return stage[i - wb]; };
                    uint32_t lit, lit_pos;
                    const uint32_t lim = (avail < kL4Stage ? avail : kL4Stage) + wb;  // staged bytes end here
                    int r = l4_literal_len(rd, wb + p, lim, kL4MaxExt, lit, lit_pos);
                    if (r == (int)kL4Bad) r = lim < in_size ? (int)kL4Esc : (int)kL4Bad;  // ran off the STAGE, not the input
                    if (r != 0) {
                        code = (uint32_t)r;
                    } else {
                        const uint32_t lit_end = lit_pos + lit;  // absolute; literals themselves are not read here
                        if (lit_end > in_size) code = kL4Bad;
                        else if (lit_end == in_size) code = kL4Last;
                        else if (in_size - lit_end < 2u) code = kL4Bad;  // no room for the offset
                        else {
                            uint32_t q = lit_end + 2u;
                            code = 0u;
                            if ((stage[p] & 15u) == 15u) {  // match-length extension bytes
                                uint32_t n = 0u, b = 255u;
                                while (b == 255u) {
                                    if (q >= in_size) { code = kL4Bad; break; }
                                    if (n == kL4MaxExt || q >= lim) { code = kL4Esc; break; }
                                    b = stage[q - wb];
                                    ++q;
                                    ++n;
                                }
                            }
                            if (code == 0u) {
                                const uint32_t rel = q - wb;  // successor relative to the window start
                                code = rel < 0xFFF0u ? rel : kL4Esc;
                            }
                        }
                    }
                }
                nx0[p] = (uint16_t)code;
                f[p] = (uint8_t)(code < kL4Win ? code : p);  // in-window successor, else a fixed point
            }
            __syncwarp();
            for (uint32_t L = 1; L < kL4Levels; ++L) {
                const uint8_t* a = f + (L - 1u) * kL4Win;
#pragma unroll
                for (uint32_t j = 0; j < 8u; ++j) {
                    const uint32_t p = lane + 32u * j;
                    f[L * kL4Win + p] = a[a[p]];
                }
                __syncwarp();
            }
            {
                const uint8_t* a = f + (kL4Levels - 1u) * kL4Win;
#pragma unroll
                for (uint32_t j = 0; j < 8u; ++j) {
                    const uint32_t p = lane + 32u * j;
                    // 2^6 + 2^6 steps: a[a[p]] is the 128th successor = the last start inside the window
                    exitp[p] = nx0[a[a[p]]];
                }
            }
        }
        __syncthreads();
        L4P(0);  // A1: successor tables + exit maps
        // ---- 2. one thread chains the true entry points through the eight exit maps ---------------------
        if (tid == 0) {
            uint32_t e = 0u, w = 0u;
            sh->stop_code = 0u;
            for (uint32_t i = 0; i < (uint32_t)kL4Warps; ++i) sh->entry[i] = 0xFFFFFFFFu;
            uint32_t next = B;
            while (w < (uint32_t)kL4Warps && B + w * kL4Win < in_size) {
                sh->entry[w] = e;
                const uint16_t* ex = reinterpret_cast<const uint16_t*>(parse_smem + w * kL4ParsePerWarp + 2u * kL4Win +
                                                                        kL4Levels * kL4Win);
                const uint32_t x = ex[e];
                if (x >= 0xFFF0u) {  // the chain of this window ends in a special sequence: at its last start
                    const uint8_t* a = reinterpret_cast<const uint8_t*>(ex) - kL4Win;  // f[levels - 1]
                    sh->stop_code = x;
                    sh->stop_pos = B + w * kL4Win + a[a[e]];
                    break;
                }
                // x >= 256: position B + w * 256 + x
                const uint32_t adv = x >> 8;
                e = x & 255u;
                w += adv;
                next = B + w * kL4Win + e;
            }
            if (sh->stop_code == 0u) sh->next_pos = next;  // chain left the super-step (or the input: checked below)
        }
        __syncthreads();
        L4P(1);  // A2: chain
        // ---- 3. per warp: enumerate the real sequence starts of its window, lengths, local scan -------
        uint32_t my_cnt = 0u, my_out = 0u;
        uint32_t lp3[3], lo3[3], excl3[3];  // first literal (window-relative), literals | offset << 16, output position
        bool val3[3];
        const uint32_t e_w = sh->entry[warp];
        if (have && e_w != 0xFFFFFFFFu) {
            uint32_t carry = 0u;
            uint32_t prev_last = 0xFFFFFFFFu;  // q of the last lane of the previous round
#pragma unroll
            for (uint32_t r = 0; r < 3u; ++r) {
                const uint32_t k = lane + 32u * r;
                uint32_t p = e_w;
#pragma unroll
                for (uint32_t L = 0; L < kL4Levels; ++L)
                    if (k & (1u << L)) p = f[L * kL4Win + p];
                // k-th successor; it stays on the last start once the chain has left the window
                uint32_t before = __shfl_up_sync(kFull, p, 1);
                if (lane == 0) before = prev_last;
                const uint32_t code = nx0[p];
                // a real, ordinary sequence: first of the window or different from its predecessor, and
                // not the special one that ends the chain (that one is handled by thread 0 below)
                const bool valid = (k == 0u || p != before) && code < 0xFFF0u;
                prev_last = __shfl_sync(kFull, p, 31);
                uint32_t olen = 0u, lpos = 0u, litoff = 0u;
                if (valid) {
                    const uint32_t tok = stage[p];
                    uint32_t lit = tok >> 4, qq = p + 1u;
                    if (lit == 15u) {
                        uint32_t b;
                        do {
                            b = stage[qq++];
                            lit += b;
                        } while (b == 255u);
                    }
                    // the offset: behind a long run of literals it may lie beyond the staged bytes (it is inside
                    // the input: A1 checked that); the match-length extension is staged (code < 0xFFF0 guarantees it)
                    uint32_t mq = qq + lit;
                    const uint32_t off = warp * kL4Win + mq + 2u <= kL4Warps * kL4Win + 32u
                                             ? (uint32_t)stage[mq] | ((uint32_t)stage[mq + 1u] << 8)
                                             : (uint32_t)in[wb + mq] | ((uint32_t)in[wb + mq + 1u] << 8);
                    mq += 2u;
                    uint32_t ml = (tok & 15u) + 4u;
                    if ((tok & 15u) == 15u) {
                        uint32_t b;
                        do {
                            b = stage[mq++];
                            ml += b;
                        } while (b == 255u);
                    }
                    olen = lit + ml;
                    lpos = qq;
                    litoff = lit | (off << 16);  // lit <= 15 + 4 * 255 here
                }
                uint32_t incl = olen;
#pragma unroll
                for (uint32_t s = 1u; s < 32u; s <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, s);
                    if (lane >= s) incl += t;
                }
                lp3[r] = lpos;
                lo3[r] = litoff;
                excl3[r] = carry + incl - olen;
                val3[r] = valid;
                carry += __shfl_sync(kFull, incl, 31);
                my_cnt += __popc(__ballot_sync(kFull, valid));
            }
            my_out = carry;
        } else {
#pragma unroll
            for (uint32_t r = 0; r < 3u; ++r) {
                lp3[r] = 0u; lo3[r] = 0u; excl3[r] = 0u; val3[r] = false;
            }
        }
        if (lane == 0) {
            sh->cnt[warp] = my_cnt;
            sh->outlen[warp] = my_out;
        }
        __syncthreads();
        L4P(2);  // A3: enumerate + parse + scan
        // ---- 4. descriptors; the special sequence that ended the chain, if any ---------------------------
        {
            uint32_t base_idx = sh->nseq, base_out = sh->out;
            for (uint32_t w = 0; w < warp; ++w) {
                base_idx += sh->cnt[w];
                base_out += sh->outlen[w];
            }
            uint32_t tot_idx = base_idx, tot_out = base_out;
            for (uint32_t w = warp; w < (uint32_t)kL4Warps; ++w) {
                tot_idx += sh->cnt[w];
                tot_out += sh->outlen[w];
            }
            if (tot_idx + 2u > desc_cap || tot_out > out_cap) {
                if (tid == 0) sh->err = -4;  // more sequences than the input can hold / output overrun
            } else {
                uint32_t rank = 0u;
#pragma unroll
                for (uint32_t r = 0; r < 3u; ++r) {
                    const uint32_t m = __ballot_sync(kFull, val3[r]);
                    if (val3[r]) {
                        const uint32_t i = base_idx + rank + __popc(m & ((1u << lane) - 1u));
                        desc[i] = L4Desc{base_out + excl3[r], wb + lp3[r], lo3[r] & 0xFFFFu, lo3[r] >> 16};
                    }
                    rank += __popc(m);
                }
            }
            __syncthreads();
            if (tid == 0 && sh->err == 0) {
                uint32_t nseq = tot_idx, out = tot_out;
                // the sequence that stopped the chain: its token position is the last start of the
                // stopping window's chain
                if (sh->stop_code != 0u) {
                    const uint32_t p = sh->stop_pos;
                    if (sh->stop_code == kL4Bad) {
                        sh->err = -1;
                    } else {
                        // serial parse from global memory (any length of extension runs)
                        auto rd = [&](uint32_t i) -> uint32_t { return in[i]; };
                        uint32_t lit, lit_pos;
                        const int r = l4_literal_len(rd, p, in_size, 0xFFFFFFFFu, lit, lit_pos);
                        if (r != 0 || lit > in_size - lit_pos) {
                            sh->err = -2;
                        } else if (lit_pos + lit == in_size) {  // last sequence: literals only
                            if (lit > out_cap - out) sh->err = -2;
                            else {
                                desc[nseq++] = L4Desc{out, lit_pos, lit, 0u};
                                out += lit;
                                sh->done = 1u;
                            }
                        } else if (in_size - (lit_pos + lit) < 2u) {
                            sh->err = -3;
                        } else {
                            uint32_t qq = lit_pos + lit + 2u;
                            const uint32_t off = (uint32_t)in[qq - 2u] | ((uint32_t)in[qq - 1u] << 8);
                            uint32_t ml = (in[p] & 15u);
                            if (ml == 15u) {
                                uint32_t b = 255u;
                                while (b == 255u) {
                                    if (qq >= in_size) { sh->err = -1; break; }
                                    b = in[qq++];
                                    ml += b;
                                    if (ml > out_cap) { sh->err = -4; break; }
                                }
                            }
                            ml += 4u;
                            if (sh->err == 0) {
                                if (lit > out_cap - out || ml > out_cap - out - lit) sh->err = -4;
                                else {
                                    desc[nseq++] = L4Desc{out, lit_pos, lit, off == 0u ? 0x10000u : off};  // (0 is invalid: B1 rejects it)
                                    out += lit + ml;
                                    sh->next_pos = qq;
                                    if (qq >= in_size) sh->err = -1;  // a match cannot be the end of a block
                                }
                            }
                        }
                    }
                } else if (sh->next_pos >= in_size) {
                    sh->err = -1;  // the chain ran off the end without a last sequence
                }
                sh->nseq = nseq;
                sh->out = out;
                sh->pos = sh->next_pos;
            }
            __syncthreads();
            L4P(3);  // A4: descriptors + special sequence
        }
    }
    if (sh->err) return sh->err;
    *total_out = sh->out;
    return (int)sh->nseq;
}

// position of a byte of the output in the shared-memory ring (v = out_pos + ga, see below)
__device__ __forceinline__ uint32_t l4_ring(uint32_t v) { return v & (kL4Ring - 1u); }

// One match byte range [a, b) of a match that starts at m with offset off (all v-space, clipped to the
// tile [tlo, ...)), bytes a + first, a + first + step, ...: a source before the tile is final already (copied
// from the history ring, or from global memory beyond its reach), a source inside the tile becomes
// the byte's parent.  An overlapping match (off < length) repeats its first `off` bytes: the parent is
// taken inside that first period, so that its chain is one hop long whatever the length.
__device__ __forceinline__ void l4_match_bytes(uint32_t a, uint32_t b, uint32_t m, uint32_t off, bool overlap,
                                               uint32_t first, uint32_t step, uint32_t tlo, uint32_t ga,
                                               uint8_t* ring, uint16_t* P, const uint8_t* out)
{
    // (out is written by this CTA: plain loads, never the read-only path)
    const uint32_t src0 = m - off;
    uint32_t v = a + first;
    if (v >= b) return;
    uint32_t r = v - m, rstep = step;  // r = (v - m) mod off, kept incrementally
    if (overlap) {
        r %= off;
        rstep %= off;
    }
    for (; v < b; v += step) {
        const uint32_t sp = src0 + r;
        if (sp < tlo) {
            ring[l4_ring(v)] = (tlo - sp <= kL4Ring - kL4Tile) ? ring[l4_ring(sp)] : out[sp - ga];
        } else {
            P[v - tlo] = (uint16_t)(sp - tlo);
        }
        r += rstep;
        if (overlap && r >= off) r -= off;
    }
}

// B3 + B4 for the 16 bytes a thread owns: EIGHT PAIRS, pair j at tile offset j * 1024 + 2 * tid.  (The first
// version gave a thread 16 consecutive bytes: for a given byte of the chunk the 32 lanes of a warp then read
// P[] 32 bytes apart -- four banks, eight-way conflicts on both hops of every round.  With pairs interleaved
// over the threads the lanes of a warp read neighbouring entries.)  The 16 parents stay in registers as packed
// u16 pairs; pointer jumping, two hops per round, a PAIR at a time -- a pair whose parents do not move any more
// sits on roots (P[x] < x for every byte that is not one) and drops out; a round reads P[], then (behind a
// barrier) every owner stores its entries --; then root -> byte.  rb[j] receives pair j's two final bytes.
#ifndef FSB_L4_HOPS
#define FSB_L4_HOPS 2  // hops per pointer-jumping round (3: A/B builds, tools/gpu_r8a.sh)
#endif
#ifndef FSB_L4_EARLY_ROOT
#define FSB_L4_EARLY_ROOT 0  // 1: a pair stops as soon as it knows that it sits on roots (A/B: tools/gpu_r11a.sh)
#endif
constexpr uint32_t kL4Pairs = kL4ByteChunk / 2u;                 // 8
constexpr uint32_t kL4PairStride = 2u * kL4Threads;              // 1024 bytes between a thread's pairs
__device__ __forceinline__ void l4_resolve(uint16_t* P, const uint8_t* ring, uint32_t tlo, uint32_t tid, uint32_t (&rb)[kL4Pairs])
{
    L4P_DECL;
    const uint32_t x0 = 2u * tid;  // tile offset of pair 0
    uint32_t pp[kL4Pairs];
    uint32_t act = 0u;  // bit j: a byte of pair j copies a byte of this tile that is not known yet
#pragma unroll
    for (uint32_t j = 0; j < kL4Pairs; ++j) {
        const uint32_t x = x0 + j * kL4PairStride;
        pp[j] = *reinterpret_cast<const uint32_t*>(P + x);
        if (pp[j] != (x | ((x + 1u) << 16))) act |= 1u << j;
    }
    // parents of the two bytes p0 = w & 0xFFFF, p1 = w >> 16: two 16-bit look-ups, no case distinction.  (One
    // 32-bit look-up when the two are an aligned pair themselves was measured: slower by 8 - 12 % -- a pair
    // stops being aligned as soon as its chain passes one odd-aligned match, the lanes of a warp then
    // disagree and the warp pays for the test AND both paths; profiles/r9g_lz4_hop_ab.txt.)
    auto hop = [&](uint32_t w) -> uint32_t { return (uint32_t)P[w & 0xFFFFu] | ((uint32_t)P[w >> 16] << 16); };
    for (;;) {
        L4P_COUNT(15, 1);  // rounds
        uint32_t changed = 0u;  // bit j: pair j has new parents
#pragma unroll
        for (uint32_t j = 0; j < kL4Pairs; ++j) {
            if (act & (1u << j)) {
#if FSB_L4_EARLY_ROOT
                // A pair leaves as soon as it KNOWS its parents are roots: after the first hop if they did not
                // move (roots already), after the second if the first hop's result did not move (the parents'
                // parents are roots: final once stored).  Without this every pair spends one more whole round
                // -- two hops -- on finding out that nothing changes any more.
                const uint32_t h1 = hop(pp[j]);
                if (h1 == pp[j]) {
                    act &= ~(1u << j);
                } else {
                    uint32_t h2 = hop(h1);
                    bool fin = h2 == h1;
#if FSB_L4_HOPS == 3
                    if (!fin) {
                        const uint32_t h3 = hop(h2);
                        fin = h3 == h2;
                        h2 = h3;
                    }
#endif
                    pp[j] = h2;
                    changed |= 1u << j;
                    if (fin) act &= ~(1u << j);
                }
#else
                uint32_t q = hop(hop(pp[j]));
#if FSB_L4_HOPS == 3
                q = hop(q);
#endif
                if (q != pp[j]) {
                    pp[j] = q;
                    changed |= 1u << j;
                } else {
                    act &= ~(1u << j);  // both parents are roots
                }
#endif
            }
        }
        __syncthreads();  // every thread has read what it needs from P[]: owners may store now (no data race)
#pragma unroll
        for (uint32_t j = 0; j < kL4Pairs; ++j)
            if (changed & (1u << j)) *reinterpret_cast<uint32_t*>(P + x0 + j * kL4PairStride) = pp[j];
#if FSB_L4_EARLY_ROOT
        if (!__syncthreads_or((int)act)) break;  // (a round without a change leaves nobody active)
#else
        if (!__syncthreads_or((int)changed)) break;
#endif
    }
    L4P(7);  // B3: pointer jumping
    // root -> byte.  Roots are literal / history bytes: final since B1.
#pragma unroll
    for (uint32_t j = 0; j < kL4Pairs; ++j) {
        const uint32_t x = x0 + j * kL4PairStride;
        uint32_t two = *reinterpret_cast<const uint16_t*>(ring + l4_ring(tlo + x));
        const uint32_t p0 = pp[j] & 0xFFFFu, p1 = pp[j] >> 16;
        if (p0 != x) two = (two & 0xFF00u) | ring[l4_ring(tlo + p0)];
        if (p1 != x + 1u) two = (two & 0x00FFu) | ((uint32_t)ring[l4_ring(tlo + p1)] << 8);
        rb[j] = two;
    }
    L4P(8);  // B4: root -> byte
}

// Phase B for one block: nseq descriptors -> out[0, total).  Returns total or a negative error.
// max_off: the largest offset the format allows (LZ4: 65535; the Zstd frames of zstd_block.cuh, whose
// sequences come through here too: the frame).
// (desc and tile_first were written by this CTA: no const / __restrict__, so that they are never read through
// the non-coherent path, which may hold the previous block's lines)
__device__ int l4_copy(const uint8_t* __restrict__ in, uint32_t in_size, uint8_t* out, uint32_t total,
                       L4Desc* desc, uint32_t nseq, uint32_t* tile_first,
                       uint8_t* ring, uint16_t* P, L4Shared* sh, L4Long* longs, L4Desc* dsm, uint32_t max_off)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // v-space: v = out_pos + ga, so that 16-byte chunks of the ring and of global memory line up
    const uint32_t ga = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
    const uint32_t vend = total + ga;
    const uint32_t ntiles = (vend + kL4Tile - 1u) / kL4Tile;
    // tile_first[t] = first descriptor whose v position is >= t * tile, nseq if there is none
    // (one pass over the descriptors; positions are a running sum, hence sorted)
    for (uint32_t j = tid; j <= nseq; j += kL4Threads) {
        const uint32_t t0 = j == 0u ? 0u : (desc[j - 1u].out_pos + ga) / kL4Tile + 1u;
        const uint32_t t1 = j < nseq ? (desc[j].out_pos + ga) / kL4Tile : ntiles;
        for (uint32_t t = t0; t <= t1 && t <= ntiles; ++t) tile_first[t] = j;
    }
    __syncthreads();
    if (sh->err) return sh->err;
    L4P_DECL;
    L4P(4);  // B0: tile index (includes nothing else)

    // descriptors of a tile: staged in shared memory by cp.async while the tile before is being resolved
    auto tile_range = [&](uint32_t t, uint32_t& j0, uint32_t& nd) {
        const uint32_t j1 = tile_first[t + 1u];  // descriptors < j1 start before the next tile
        j0 = tile_first[t];
        if (j0 > 0u) --j0;                       // the one before may reach into the tile
        nd = j1 - j0;
    };
    auto stage_descs = [&](uint32_t t) {
        uint32_t j0, nd;
        tile_range(t, j0, nd);
        const uint32_t ns = nd + 1u < kL4DescStage ? nd + 1u : kL4DescStage;  // one more: where the last one ends
        for (uint32_t k = tid; k < ns; k += kL4Threads) {
            if (j0 + k < nseq) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(desc + j0 + k) : "memory");
            } else {
                dsm[k] = L4Desc{total, in_size, 0u, 0u};
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (ntiles != 0u) stage_descs(0u);

    for (uint32_t t = 0; t < ntiles; ++t) {
        L4P_COUNT(13, 1);  // tiles
        const uint32_t tlo = t * kL4Tile;                                   // v-space
        const uint32_t thi = (tlo + kL4Tile < vend) ? tlo + kL4Tile : vend;
        uint32_t j0, nd;
        tile_range(t, j0, nd);
        const uint32_t ns = nd + 1u < kL4DescStage ? nd + 1u : kL4DescStage;
        const uint32_t lo_v = t == 0u ? ga : tlo;               // the first tile starts ga bytes in
        // every byte of the tile starts as its own root
        {
            const uint32_t e0 = tid * kL4ByteChunk;
            uint32_t id[8];
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j) id[j] = (e0 + 2u * j) | ((e0 + 2u * j + 1u) << 16);
            *reinterpret_cast<uint4*>(P + e0) = make_uint4(id[0], id[1], id[2], id[3]);
            *reinterpret_cast<uint4*>(P + e0 + 8u) = make_uint4(id[4], id[5], id[6], id[7]);
        }
        if (tid == 0) sh->n_long = 0u;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        L4P(6);  // B1a: identity parents, waiting for the descriptors

        // ---- B1: one thread per sequence: literals -> ring, match bytes -> parents (or history bytes -> ring)
        for (uint32_t k = tid; k < nd; k += kL4Threads) {
            uint4 d;
            uint32_t oe;
            if (k + 1u < ns) {
                d = *reinterpret_cast<const uint4*>(dsm + k);
                oe = dsm[k + 1u].out_pos;
            } else {  // (more sequences in the tile than the stage holds: incompressible input)
                d = *reinterpret_cast<const uint4*>(desc + j0 + k);
                oe = j0 + k + 1u < nseq ? desc[j0 + k + 1u].out_pos : total;
            }
            const uint32_t o = d.x + ga;
            oe += ga;
            if (oe <= lo_v || o >= thi) continue;  // (the descriptor in front of the tile may end before it)
            const uint32_t lit_pos = d.y, lit = d.z, off = d.w;
            if (oe < o || lit_pos > in_size || lit > in_size - lit_pos || lit > oe - o) {
                atomicCAS(&sh->err, 0, -2);
                break;
            }
            const uint32_t m = o + lit;   // first match byte
            const uint32_t ml = oe - m;   // 0: the block's last sequence
            // literals inside the tile: a few bytes at most for FLAG data (incompressible input: one long run)
            if (lit != 0u) {
                const uint32_t a = o > lo_v ? o : lo_v, b = m < thi ? m : thi;
                if (a < b && b - a > kL4LongMatch) {
                    const uint32_t i = atomicAdd(&sh->n_long, 1u);
                    if (i < kL4MaxLong) longs[i] = L4Long{a, b, lit_pos - o, 0u};
                } else {
                    for (uint32_t v = a; v < b; ++v) ring[l4_ring(v)] = in[lit_pos + (v - o)];
                }
            }
            if (ml == 0u) continue;
            if (off == 0u || off > max_off || off > m - ga) {
                atomicCAS(&sh->err, 0, -4);
                break;
            }
            const uint32_t a = m > lo_v ? m : lo_v, b = oe < thi ? oe : thi;
            if (a >= b) continue;
            if (b - a > kL4LongMatch) {
                const uint32_t i = atomicAdd(&sh->n_long, 1u);
                if (i < kL4MaxLong) longs[i] = L4Long{a, b, m, off};  // (always: a tile has room for no more)
                continue;
            }
            uint32_t x = a - tlo;
            const uint32_t y = b - tlo;
            if (off >= ml && a - off < tlo && tlo + off - a <= kL4Ring - kL4Tile) {
                // a plain copy that begins in the history ring (an earlier tile: final bytes): copy those,
                // two at a time where FLAG words line up; what is left of the match has its source in this tile
                const uint32_t he = b < tlo + off ? b : tlo + off;  // v < tlo + off <=> source before the tile
                uint32_t v = a;
                if (((v | off) & 1u) == 0u)
                    for (; v + 2u <= he; v += 2u)
                        *reinterpret_cast<uint16_t*>(ring + l4_ring(v)) = *reinterpret_cast<const uint16_t*>(ring + l4_ring(v - off));
                for (; v < he; ++v) ring[l4_ring(v)] = ring[l4_ring(v - off)];
                x = he - tlo;
            }
            if (x >= y) continue;
            const bool wraps = off < ml;  // a repeating match: a run of equal FLAG words is offset 2
            if (wraps ? (m - off >= tlo && a == m && (off == 1u || ((off | x) & 1u) == 0u)) : x >= off) {
                // The common cases, ONE path for both (the lanes of a warp hold a mix of them): a plain copy
                // whose source lies inside this tile, parent(v) = v - off, and a repeating match whose first
                // period lies inside it, parent(v) = first period + (v - m) mod off -- its chain is one hop long
                // whatever the length.  Byte d = v - m of the match has parent s0 + (d mod period); the period of
                // a plain copy is never reached.  Two parents per 32-bit store.
                const uint32_t s0 = m - off - tlo;          // (as a 32-bit sum with r: never negative)
                uint32_t r = x + tlo - m;                   // d of the first byte (0 for a repeating match)
                const uint32_t lim = wraps ? off : 0xFFFFFFFFu;
                const uint32_t inc = (wraps && off == 1u) ? 0u : 1u;  // offset 1: every byte copies the same one
                if (x & 1u) {
                    P[x] = (uint16_t)(s0 + r);
                    ++x;
                    r += inc;
                }
                for (; x + 2u <= y; x += 2u) {
                    const uint32_t par = s0 + r;
                    *reinterpret_cast<uint32_t*>(P + x) = (par & 0xFFFFu) | ((par + inc) << 16);
                    r += 2u * inc;
                    if (r >= lim) r -= lim;
                }
                if (x < y) P[x] = (uint16_t)(s0 + r);
            } else {
                l4_match_bytes(a, b, m, off, off < ml, 0u, 1u, tlo, ga, ring, P, out);
            }
        }
        __syncthreads();
        L4P(5);  // B1c: sequences
        if (sh->err) return sh->err;
        if (t + 1u < ntiles) stage_descs(t + 1u);  // (nobody reads this tile's any more; they land during B3)
        // long matches and long runs of literals: one warp each, 32 bytes per step
        {
            const uint32_t nl = sh->n_long < kL4MaxLong ? sh->n_long : kL4MaxLong;
            for (uint32_t i = warp; i < nl; i += (uint32_t)kL4Warps) {
                const L4Long g = longs[i];
                if (g.off == 0u) {
                    for (uint32_t v = g.a + lane; v < g.b; v += 32u) ring[l4_ring(v)] = in[v + g.m];
                } else {
                    l4_match_bytes(g.a, g.b, g.m, g.off, g.off < g.b - g.m, lane, 32u, tlo, ga, ring, P, out);
                }
            }
            if (nl != 0u) __syncthreads();
        }
        L4P(14);  // B1d: long matches
        // ---- B3, B4: parents -> roots -> bytes
        uint32_t rb[kL4Pairs];
        l4_resolve(P, ring, tlo, tid, rb);
        L4P_RESET;
        __syncthreads();  // every thread has read the roots it needs: the tile's ring bytes may be overwritten now
        // ---- B5: the pairs go to the ring (history of the next tiles) and to global memory straight from the
        // registers: a warp stores 64 consecutive bytes at a time
#pragma unroll
        for (uint32_t j = 0; j < kL4Pairs; ++j) {
            const uint32_t v = tlo + 2u * tid + j * kL4PairStride;
            *reinterpret_cast<uint16_t*>(ring + l4_ring(v)) = (uint16_t)rb[j];
            if (v >= lo_v && v + 2u <= thi) {
                *reinterpret_cast<uint16_t*>(out + (v - ga)) = (uint16_t)rb[j];  // (out - ga is 16-byte aligned, v even)
            } else {
                if (v >= lo_v && v < thi) out[v - ga] = (uint8_t)rb[j];
                if (v + 1u >= lo_v && v + 1u < thi) out[v + 1u - ga] = (uint8_t)(rb[j] >> 8);
            }
        }
        __syncthreads();  // P[] and the chunk's ring bytes are free for the next tile
        L4P(9);  // B5: chunk -> ring + global
    }
    return (int)total;
}

// status[b] = decoded size (must equal raw_size) or a negative error code.  scratch: gridDim.x
// regions of scratch_stride bytes (l4_scratch_bytes of the largest block).
__global__ void __launch_bounds__(kL4Threads, 2)  // two CTAs per SM: at most 64 registers
lz4_decode_cta_kernel(const uint8_t* __restrict__ comp, uint8_t* raw, const Lz4BlockDesc* __restrict__ bdesc,
                      int* __restrict__ status, uint32_t n_blocks, unsigned char* scratch, size_t scratch_stride,
                      uint32_t desc_cap)
{
    extern __shared__ __align__(16) unsigned char l4_smem[];
    uint8_t* ring = l4_smem;
    uint16_t* P = reinterpret_cast<uint16_t*>(l4_smem + kL4Ring);
    L4Shared* sh = reinterpret_cast<L4Shared*>(l4_smem + kL4Ring + 2u * kL4Tile);
    L4Long* longs = reinterpret_cast<L4Long*>(l4_smem + kL4Ring + 2u * kL4Tile + 256u);
    L4Desc* dsm = reinterpret_cast<L4Desc*>(l4_smem + kL4Ring + 2u * kL4Tile + 256u + sizeof(L4Long) * kL4MaxLong);
    L4Desc* desc = reinterpret_cast<L4Desc*>(scratch + (size_t)blockIdx.x * scratch_stride);
    uint32_t* tile_first = reinterpret_cast<uint32_t*>(desc + desc_cap);

    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const Lz4BlockDesc d = bdesc[b];
        if (threadIdx.x == 0) sh->err = 0;
        __syncthreads();
        int r = 0;
        if (d.comp_size != 0u) {
            uint32_t total = 0u;
            const int nseq = l4_parse(comp + d.comp_off, d.comp_size, d.raw_size, desc, desc_cap, ring, sh, &total);
            __syncthreads();
            r = nseq;
            if (nseq >= 0)
                r = l4_copy(comp + d.comp_off, d.comp_size, raw + d.raw_off, total, desc, (uint32_t)nseq, tile_first,
                            ring, P, sh, longs, dsm, 0xFFFFu);
        }
        __syncthreads();
        if (threadIdx.x == 0) status[b] = r;
        __syncthreads();
    }
}

}  // namespace fsb200
