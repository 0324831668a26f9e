// flagstat_kernel_dyn.cuh -- the group kernel with its work handed out DYNAMICALLY, per warp.
//
// Why: the per-CTA timeline of the statically strided group kernel (tools/timeline_probe.py,
// profiles/r4e_timeline.jsonl) shows its CTAs leaving the main loop anywhere between 194 and 240 us
// on the 1.65 GB column (median 217) -- two CTAs sharing an SM do not progress alike, and some SMs
// are persistently 10 % slower than others -- so a launch lasts as long as its slowest CTA while
// the HBM idles behind the fast ones.  At 100 M records the spread is 21 ... 36 us around a
// median of 28.  A static split cannot fix that; a shared work counter can.
//
// How: the column is cut into CHUNKS of CG groups x 4 batches x 2 KiB = CG x 8 KiB, the bytes ONE
// WARP consumes per CG laps of its depth-4 cp.async ring (lane l reads 16 bytes at
// chunk + (4 * batch + u) * 512 + 16 * l: every load instruction is one contiguous 512-byte
// request, as in the static kernel).  Warp w of every CTA draws from counter w of the launch's
// slot: chunk 8 * c + w for the c-th claim, so the eight counters sweep the column together as
// one compact front.  Eight counters because ONE cannot keep up: same-address atomics top out
// near 0.36 per ns on this part (measured: a single counter with 8 KiB chunks took the HiSeqX
// column from 236 to 557 us, profiles/r4g_length_sweep_quick.jsonl), and 7 TB/s in 8 KiB claims
// is 0.85 per ns -- 0.11 per counter.  Every SM has one warp on every counter, so a slow SM
// claims less from all eight and the counters run dry together.  Claims are made ONE CHUNK
// AHEAD (the claim issued at the top of a chunk is first needed when the ring is refilled for
// the chunk after the next, >= 2 us later), and a warp's FIRST chunk is static (8 * blockIdx + w):
// the ring is primed without waiting for anything and counter w starts handing out claim
// gridDim.x.  No CTA-wide synchronisation, no shared-memory hand-off.
//
// The slot (8 counters on separate 128-byte lines, done tickets, owner) comes from a per-device
// table (flagstat_capi.cu, dyn_slot): every stream gets a PAIR of slots and alternates between
// them, so a launch with the programmatic-serialization attribute starts on its own counters
// while its predecessor drains.  A third launch cannot be live together with the first:
// every CTA of the second waits in its epilogue for the first to complete (griddepcontrol.wait)
// and holds its SM slot meanwhile, every launch of this kernel asks for exactly as many CTAs as
// an empty machine can hold, so "all CTAs of the second have started" -- the condition for the
// third to start -- cannot be true while a CTA of the first is still resident.  The warp that
// draws the last done-ticket zeroes the slot and releases it.  As a second line of defence
// (a destroyed stream whose handle value is reused while its work is still running) the slot
// carries an owner tag: lane 0 of every warp checks it before its first claim, installs its own
// on a free slot, and waits while another launch holds it.  Both round trips hide behind the
// loads of the first chunk.  Launches under stream capture (a graph would bake the slot in and
// concurrent replays would share it), on cudaStreamPerThread (one handle, many streams) and
// columns too short to fill the grid run the static kernel.
//
// Left-overs (< one chunk of vectors, plus the <= 14 ragged records around the 16-byte
// aligned body) are taken by warp 0 of CTA 0 before its loop, as in the static kernel.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "flagstat_kernel_group.cuh"

namespace fsb200 {

constexpr int kDynWarpVecPerBatch = 32 * kU;             // 128 uint4 = 2 KiB per warp batch
constexpr int kDynWarpVecPerGroup = 4 * kDynWarpVecPerBatch;  // 8 KiB
// one slot: kWarps counters on separate 128-byte lines, then {done tickets, owner tag}
constexpr unsigned kDynCtrStride = 32;                                   // in 4-byte words
constexpr unsigned kDynSlotBytes = (kWarps + 1) * kDynCtrStride * 4;     // 1152

template <int MODE, int VARIANT, int MINB, int CG>
__global__ void __launch_bounds__(kThreads, MINB)
flagstat_kernel_dyn(const uint16_t* __restrict__ base, uint64_t n,
                    unsigned long long* __restrict__ out, const __grid_constant__ XchgArgs xa)
{
    constexpr int DEPTH = 4;
    constexpr uint32_t kChunkVec = (uint32_t)CG * kDynWarpVecPerGroup;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (xa.pdl) pdl_launch_dependents();
    FSB_TL(0);

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t addr = reinterpret_cast<uint64_t>(base);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint4* __restrict__ body = reinterpret_cast<const uint4*>(base + head);
    const uint64_t V = (n - head) >> 3;
    const uint64_t tail_start = head + (V << 3);
    const uint64_t ntail = n - tail_start;
    const uint32_t NCH = (uint32_t)(V / kChunkVec);  // full chunks (the host keeps V / kChunkVec < 2^32)
    unsigned int* const slot = reinterpret_cast<unsigned int*>(xa.dyn);
    unsigned int* const ctr = slot + warp * kDynCtrStride;          // this warp index's counter
    unsigned int* const done = slot + kWarps * kDynCtrStride;       // done tickets, [1] = owner tag
    const uint32_t W = gridDim.x * (uint32_t)kWarps;
    const uint32_t G = gridDim.x;

    GroupLanes<MODE, VARIANT> st;
    st.clear();
    uint32_t groups = 0;  // groups absorbed this epoch (warp-uniform)
    unsigned long long acc_all = 0ull, acc_fail = 0ull;

    // this lane's 16 bytes of load u of stage s:  smem + ((s * kU + u) * kThreads + tid) * 16
    const uint32_t my_smem = smem_u32(smem_raw) + tid * 16u;
    const unsigned char* const body_lane = reinterpret_cast<const unsigned char*>(body) + lane * 16u;
    auto chunk_ptr = [&](uint32_t c) -> const unsigned char* {
        return body_lane + (uint64_t)c * (kChunkVec * 16u);
    };
    // refill stage S with batch S of the group at `src` (or commit an empty group)
    auto fetch = [&](auto stage_tag, const unsigned char* src, bool valid) {
        constexpr int S = decltype(stage_tag)::value;
        if (valid) {
            cp_async16_imm<(S * kU + 0) * kThreads * 16, (S * kU + 0) * 512>(my_smem, src);
            cp_async16_imm<(S * kU + 1) * kThreads * 16, (S * kU + 1) * 512>(my_smem, src);
            cp_async16_imm<(S * kU + 2) * kThreads * 16, (S * kU + 2) * 512>(my_smem, src);
            cp_async16_imm<(S * kU + 3) * kThreads * 16, (S * kU + 3) * 512>(my_smem, src);
        }
        cp_async_commit();
    };
    using P0 = std::integral_constant<int, 0>;
    using P1 = std::integral_constant<int, 1>;
    using P2 = std::integral_constant<int, 2>;
    using P3 = std::integral_constant<int, 3>;

    // first chunk: static, so the ring is primed before anything has to be waited for
    uint32_t cur = blockIdx.x * (uint32_t)kWarps + warp, nxt = 0u, pend = 0u;  // claim c of counter w = chunk 8 c + w
    {
        const bool v = cur < NCH;
        const unsigned char* p = chunk_ptr(cur);
        fetch(P0{}, p, v);
        fetch(P1{}, p, v);
        fetch(P2{}, p, v);
        fetch(P3{}, p, v);
    }
    FSB_TL(1);
    // the slot is this launch's (see the header comment), then the first dynamic claim
    if (lane == 0) {
        const unsigned int tag = xa.dyn_tag;
        unsigned int o = ld_acquire_gpu_u32(done + 1);
        while (o != tag) {
            if (o == 0u) {
                o = atomicCAS(done + 1, 0u, tag);
                if (o == 0u) o = tag;
            } else {
                __nanosleep(256);
                o = ld_acquire_gpu_u32(done + 1);
            }
        }
        nxt = (G + atomicAdd(ctr, 1u)) * (uint32_t)kWarps + warp;
    }

    if (blockIdx.x == 0 && warp == 0) {
        // vectors behind the last full chunk: < kChunkVec of them = up to 4 * CG zero-padded warp batches
        const uint64_t left0 = (uint64_t)NCH * kChunkVec;
#pragma unroll 1
        for (uint32_t b = 0; left0 + (uint64_t)b * kDynWarpVecPerBatch < V; ++b) {
            uint32_t w[16];
            uint4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint64_t idx = left0 + (uint64_t)b * kDynWarpVecPerBatch + (uint64_t)u * 32u + lane;
                v[u] = (idx < V) ? ld_stream(body + idx) : make_uint4(0u, 0u, 0u, 0u);
            }
            unpack4(v, w);
            st.template step<0>(w);
            st.close();  // every left-over batch is its own partial group: no run-time POS
            ++groups;
        }
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
        if (lane < head) w[0] = base[lane];
        else if (lane - head < ntail) w[0] = base[tail_start + (lane - head)];
        st.template step<0>(w);
        st.close();
        ++groups;
    }

    nxt = __shfl_sync(0xffffffffu, nxt, 0);

    auto batch = [&](auto pos_tag, bool dense, const unsigned char* refill, bool refill_valid) -> bool {
        constexpr int POS = decltype(pos_tag)::value;
        cp_async_wait<DEPTH - 1>();
        uint32_t w[16];
        lds128_imm<(POS * kU + 0) * kThreads * 16>(my_smem, w[0], w[1], w[2], w[3]);
        lds128_imm<(POS * kU + 1) * kThreads * 16>(my_smem, w[4], w[5], w[6], w[7]);
        lds128_imm<(POS * kU + 2) * kThreads * 16>(my_smem, w[8], w[9], w[10], w[11]);
        lds128_imm<(POS * kU + 3) * kThreads * 16>(my_smem, w[12], w[13], w[14], w[15]);
        const bool general = st.template step<POS>(w, dense);
        fetch(pos_tag, refill, refill_valid);
        return general;
    };
    uint32_t dense_left = 0u;
    uint32_t gi = 0u;  // group inside the current chunk
    // One flush site: every pass of the outer loop is one epoch of at most kMaxGroups groups.
    do {
        uint32_t room = GCounter::kMaxGroups - groups;
        while (room != 0u && cur < NCH) {
            if (gi == 0u && lane == 0)  // the chunk after `nxt`; read one chunk later
                pend = (G + atomicAdd(ctr, 1u)) * (uint32_t)kWarps + warp;
            // where the ring is refilled from while this group is consumed: the next group of this
            // chunk, or the first group of the next chunk
            const bool last_of_chunk = gi + 1u == (uint32_t)CG;
            const unsigned char* rp = last_of_chunk ? chunk_ptr(nxt) : chunk_ptr(cur) + (gi + 1u) * (kDynWarpVecPerGroup * 16u);
            const bool rv = last_of_chunk ? (nxt < NCH) : true;
            const bool dense = VARIANT == 3 && dense_left != 0u;
            const bool g0 = batch(P0{}, dense, rp, rv);
            const bool g1 = batch(P1{}, dense, rp, rv);
            const bool g2 = batch(P2{}, dense, rp, rv);
            const bool g3 = batch(P3{}, dense, rp, rv);
            if (VARIANT == 3) {
                if (dense) --dense_left;
                else if (g0 && g1 && g2 && g3) dense_left = kDenseGroups;
            }
            --room;
            if (last_of_chunk) {
                gi = 0u;
                cur = nxt;
                nxt = __shfl_sync(0xffffffffu, pend, 0);
            } else {
                ++gi;
            }
        }
        FSB_TL(3);
        acc_all += st.all.flush_warp(lane);
        if (MODE != kPospopcnt && st.fail_dirty) acc_fail += st.fail.flush_warp(lane);
        st.clear();
        groups = 0;
    } while (cur < NCH);
    cp_async_wait<0>();
    FSB_TL(4);

    // the last warp of the launch to get here zeroes the slot and releases it
    {
        unsigned int t = 0u;
        if (lane == 0) t = atomicAdd(done, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t == W - 1u) {
            if (lane < (uint32_t)kWarps) atomicExch(slot + lane * kDynCtrStride, 0u);
            if (lane == 0) atomicExch(done, 0u);
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(done + 1, 0u);
        }
    }

    cta_epilogue<MODE>(out, acc_all, acc_fail, n, warp, lane, true, xa);
    FSB_TL(5);
}

}  // namespace fsb200
