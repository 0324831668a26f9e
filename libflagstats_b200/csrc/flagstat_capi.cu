// flagstat_capi.cu -- C ABI of libflagstats_cuda.so (include/flagstats_cuda.h).
//
// Host-side plumbing only: pointer classification, staging of host data over
// PCIe in overlapped chunks, the pinned block ring of the streaming API, the
// per-device contexts.  All counting happens in flagstat_kernels.cuh.
// There is deliberately NO CPU implementation behind these entry points: if no
// device is usable they return FLAGSTAT_CUDA_ENODEV.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <emmintrin.h>
#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/flagstats_cuda.h"
#include "flagstat_kernels.cuh"
#include "flagstat_kernel_tma.cuh"
#include "flagstat_kernel_group.cuh"
#ifdef FSB_ALL_VARIANTS
#include "flagstat_kernel_dyn.cuh"  // dynamically scheduled twin of the default kernel: a measured non-gain, A/B builds only
#endif
#include "synth.cuh"
#include "lz4_block.cuh"
#include "lz4_block_group.cuh"
#include "lz4_block_cta.cuh"
#include "zstd_block.cuh"
#include "ingest_text.cuh"
#include <cub/device/device_scan.cuh>

namespace {

using namespace fsb200;

constexpr int kMaxDevices = 64;
constexpr size_t kChunkBytes = 32u << 20;  // staging chunk for host data
constexpr int kStages = 3;

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_variant{-1};
std::atomic<int> g_ctas_per_sm{0};
std::atomic<uint32_t> g_min_len{0};  // 0 = not initialised

#define CK(expr)                              \
    do {                                      \
        cudaError_t e_ = (expr);              \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

// Makes `dev` the current device for a scope and puts the caller's device back afterwards, so
// that a handle bound to one GPU can be used from a thread whose current device is another.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        else prev = -1;  // nothing to restore
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// The C ABI promises that nothing throws: entries that use std::vector / std::string /
// std::thread run their body through this (allocation failure -> FLAGSTAT_CUDA_ENOMEM).
template <class F>
int guarded(F&& body) noexcept
{
    try {
        return body();
    } catch (...) {
        return FLAGSTAT_CUDA_ENOMEM;
    }
}

// RAII for the measurement aids (FLAGSTAT_cuda_time_device*, _read_probe): nothing leaks on an early CK() return
struct TimerPair {
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool own_stream = false;
    ~TimerPair()
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (own_stream && st) cudaStreamDestroy(st);
    }
};

// Kernel variants (FLAGSTAT_cuda_set_variant / env FLAGSTAT_CUDA_VARIANT).  All
// compute the same thing; they differ in how bytes reach the registers and in
// how the counters are organised, and are kept selectable for A/B measurement
// (profiles/, tools/variant_ab.py, tools/perf_sweep.py):
//   0  cp.async ring + compile-time groups of 4 batches; QC-clean batches without SECONDARY
//      records use the fp16x2-compare mask of variant 8, all others build the keep-mask and the
//      SECONDARY fix-up with HFMA2 and gate the QC-fail counter with HMUL2 on the FMA pipe;
//      detect-free dense mode while every batch needs the second counter          (default)
//   1  the group kernel with the integer-only mask select (cross-check of the fp16 forms)
//   2  register-staged LDG.128 double buffer, run-time hold levels
//   3  TMA (cp.async.bulk + mbarrier) shared-memory ring, 4 stages, producer warp
//   4  TMA ring, 6 stages
//   5  cp.async ring depth 4 with run-time hold levels (the round-1 "r1i" kernel)
//   6  same, depth 2
//   7  as 5 with the class tests as fp16x2 tent functions (half pipe instead of ALU pipe)
//   8  the group kernel with the fp16x2-compare mask + PRMT fail mask on every path (the
//      default until profiles/r2f: 0.71-0.76 of the measured peak on U(0,4095), 0.92 now)
constexpr int kNumVariants = 9;
using KernelFn = void (*)(const uint16_t*, uint64_t, unsigned long long*, const XchgArgs);

struct KernelCfg {
    KernelFn fn[2];   // [mode]
    int threads;
    size_t smem;
};

constexpr size_t tma_smem(int stages) { return (size_t)stages * kStageBytes + 2u * stages * 8u; }

// The product library ships the default (0), the integer-only cross-check (1), one TMA variant (3)
// and the compare-mask forms (8, the cross-check of the FMA-pipe mask arithmetic).  The others are
// measured A/B history (profiles/README.md): compiled only with -DFSB_ALL_VARIANTS
// (python -m libflagstats_b200.build --all-variants -> tools/bin/libflagstats_cuda_variants.so).
#ifdef FSB_ALL_VARIANTS
#define FSB_AB(...) __VA_ARGS__
#else
#define FSB_AB(...) {{nullptr, nullptr}, 0, 0}
#endif
const KernelCfg kKernels[kNumVariants] = {
    {{flagstat_kernel_group<kFlagstat, 3, 2>, flagstat_kernel_group<kPospopcnt, 0, 2>},
     kThreads, (size_t)4 * kStageBytes},
    {{flagstat_kernel_group<kFlagstat, 1, 2>, flagstat_kernel_group<kPospopcnt, 0, 2>},
     kThreads, (size_t)4 * kStageBytes},
    FSB_AB({{flagstat_kernel<kFlagstat, 0>, flagstat_kernel<kPospopcnt, 0>}, kThreads, 0}),
    {{flagstat_kernel_tma<kFlagstat, 0, 4, 2>, flagstat_kernel_tma<kPospopcnt, 0, 4, 2>},
     kThreads + 32, tma_smem(4)},
    FSB_AB({{flagstat_kernel_tma<kFlagstat, 0, 6, 2>, flagstat_kernel_tma<kPospopcnt, 0, 6, 2>},
            kThreads + 32, tma_smem(6)}),
    FSB_AB({{flagstat_kernel_ring<kFlagstat, 0, 4, 2>, flagstat_kernel_ring<kPospopcnt, 0, 4, 2>},
            kThreads, (size_t)4 * kStageBytes}),
    FSB_AB({{flagstat_kernel_ring<kFlagstat, 0, 2, 2>, flagstat_kernel_ring<kPospopcnt, 0, 2, 2>},
            kThreads, (size_t)2 * kStageBytes}),
    FSB_AB({{flagstat_kernel_ring<kFlagstat, 2, 4, 2>, flagstat_kernel_ring<kPospopcnt, 0, 4, 2>},
            kThreads, (size_t)4 * kStageBytes}),
    {{flagstat_kernel_group<kFlagstat, 0, 2>, flagstat_kernel_group<kPospopcnt, 0, 2>},
     kThreads, (size_t)4 * kStageBytes},
};

inline bool variant_built(int v) { return v >= 0 && v < kNumVariants && kKernels[v].fn[0] != nullptr; }

// kSamtools (exact n_pair_all in slots 0 / 16) has one instantiation whatever variant is
// selected: the group kernel with the FMA-pipe mask forms (flagstat_kernels.cuh: mask_select_fx)
const KernelCfg kSamtoolsKernel = {
    {flagstat_kernel_group<kSamtools, 3, 2>, flagstat_kernel_group<kSamtools, 3, 2>},
    kThreads, (size_t)4 * kStageBytes};

// The default variant's dynamically scheduled twin (flagstat_kernel_dyn.cuh), by mode: every warp
// claims 8 / 16 KiB chunks from per-launch counters instead of taking a static share.  It does
// what it was built for -- the CTAs of a launch leave the loop within 2 % of each other instead
// of 194 ... 232 us apart on the 1.65 GB column -- and the launch is no faster for it
// (244 vs 243 us sustained, 41.6 vs 40.6 us at 100 M records; profiles/r4h_*): the CTAs that a
// static split lets finish early leave their share of the HBM to the others, the machine is
// limited by its aggregate rate either way.  Kept for A/B builds (-DFSB_ALL_VARIANTS) with its
// tests; the product library does not contain it.
#ifdef FSB_ALL_VARIANTS
constexpr int kDynCG = 1;
const KernelFn kDynKernels[2][3] = {
    {flagstat_kernel_dyn<kFlagstat, 3, 2, 1>, flagstat_kernel_dyn<kPospopcnt, 0, 2, 1>,
     flagstat_kernel_dyn<kSamtools, 3, 2, 1>},
    {flagstat_kernel_dyn<kFlagstat, 3, 2, 2>, flagstat_kernel_dyn<kPospopcnt, 0, 2, 2>,
     flagstat_kernel_dyn<kSamtools, 3, 2, 2>},
};
constexpr int kDynStreams = 1024;  // streams per device that can hold a private pair of counter slots

struct DynStream {
    int base;        // slot pair index
    uint64_t count;  // launches of the dynamic kernel on this stream so far: parity picks the slot
};
#endif

struct DeviceInfo {
    int sms = 0;
    int occ[kNumVariants][2];  // resident CTAs per SM
    int occ_samtools = 2;
#ifdef FSB_ALL_VARIANTS
    int occ_dyn[3] = {2, 2, 2};
    unsigned char* dyn = nullptr;  // kDynStreams x 2 slots of kDynSlotBytes (device memory, zeroed)
    std::unordered_map<cudaStream_t, DynStream> dyn_map;
    std::mutex dyn_mu;
    std::atomic<uint32_t> dyn_tag{0};
#endif
    bool ok = false;
};

std::mutex g_mu;
int g_ndev = -2;  // -2 = not probed
DeviceInfo g_dev[kMaxDevices];

int probe_devices()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ndev != -2) return g_ndev;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (n > kMaxDevices) n = kMaxDevices;
    g_ndev = n;
    return n;
}

int device_info(int dev, DeviceInfo** out)
{
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (dev < 0 || dev >= g_ndev) return FLAGSTAT_CUDA_EINVAL;
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceInfo& d = g_dev[dev];
    if (!d.ok) {
        int cur = -1;
        CK(cudaGetDevice(&cur));
        if (cur != dev) CK(cudaSetDevice(dev));
        CK(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        for (int v = 0; v < kNumVariants; ++v)
            for (int m = 0; m < 2; ++m) {
                const KernelCfg& k = kKernels[v];
                if (!variant_built(v)) {
                    d.occ[v][m] = 1;
                    continue;
                }
                if (k.smem > 48u * 1024u)
                    CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(k.fn[m]),
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)k.smem));
                int nb = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &nb, reinterpret_cast<const void*>(k.fn[m]), k.threads, k.smem) !=
                    cudaSuccess) {
                    cudaGetLastError();
                    nb = 2;
                }
                d.occ[v][m] = nb < 1 ? 1 : nb;
            }
        {
            const KernelCfg& k = kSamtoolsKernel;
            CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(k.fn[0]),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem));
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &nb, reinterpret_cast<const void*>(k.fn[0]), k.threads, k.smem) != cudaSuccess) {
                cudaGetLastError();
                nb = 2;
            }
            d.occ_samtools = nb < 1 ? 1 : nb;
        }
        // the LZ4 block decoders of flagstat_blockfile.inl (> 48 KiB of dynamic shared memory)
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(lz4_decode_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLz4Smem));
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(lz4_decode_group_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLz4GroupSmem));
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(lz4_decode_cta_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL4Smem));
        // ... and the two kernels of the Zstd decoder (zstd_block.cuh)
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(zstd_parse_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kZstdParseSmem));
        CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(zstd_copy_kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kL4Smem));
#ifdef FSB_ALL_VARIANTS
        for (int c = 0; c < 2; ++c)
            for (int m = 0; m < 3; ++m) {
                const void* fn = reinterpret_cast<const void*>(kDynKernels[c][m]);
                CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kStageBytes));
                int nb = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, kThreads, (size_t)4 * kStageBytes) !=
                    cudaSuccess) {
                    cudaGetLastError();
                    nb = 2;
                }
                if (c == 0) d.occ_dyn[m] = nb < 1 ? 1 : nb;
            }
        CK(cudaMalloc(&d.dyn, (size_t)kDynStreams * 2 * kDynSlotBytes));
        CK(cudaMemset(d.dyn, 0, (size_t)kDynStreams * 2 * kDynSlotBytes));
#endif
        if (cur != dev && cur >= 0) CK(cudaSetDevice(cur));
        d.ok = true;
    }
    *out = &d;
    return 0;
}

#ifdef FSB_ALL_VARIANTS
// Dynamic scheduling knobs (FLAGSTAT_cuda_set_dynamic; env FLAGSTAT_CUDA_DYNAMIC=1 switches it on,
// FLAGSTAT_CUDA_DYN_MIN_CHUNKS / FLAGSTAT_CUDA_DYN_CG preset the other two):
//   g_dyn_min_chunks  -1 off (default); 0 = from 6 chunks per resident warp (~58 M records on a B200);
//                     > 0: use the dynamic kernel from that many chunks on (tests force 1)
//   g_dyn_cg          8 KiB groups per claimed chunk: 1 or 2
std::atomic<long long> g_dyn_min_chunks{-2};  // -2 = read the environment first
std::atomic<int> g_dyn_cg{0};

long long dyn_min_chunks()
{
    long long v = g_dyn_min_chunks.load(std::memory_order_relaxed);
    if (v == -2) {
        v = -1;  // off unless asked for: it is an A/B kernel
        if (const char* e = std::getenv("FLAGSTAT_CUDA_DYNAMIC"))
            if (std::atoi(e) != 0) v = 0;
        if (const char* e = std::getenv("FLAGSTAT_CUDA_DYN_MIN_CHUNKS"))
            if (std::atoll(e) > 0) v = std::atoll(e);
        g_dyn_min_chunks.store(v);
    }
    return v;
}

int dyn_cg()
{
    int v = g_dyn_cg.load(std::memory_order_relaxed);
    if (v == 0) {
        const char* e = std::getenv("FLAGSTAT_CUDA_DYN_CG");
        v = (e && std::atoi(e) == 2) ? 2 : kDynCG;
        g_dyn_cg.store(v);
    }
    return v;
}

// The counter slot of this launch and its owner tag (flagstat_kernel_dyn.cuh explains why a pair
// of slots per stream, used alternately, is private to each launch), or nullptr when the launch
// has to run the static kernel: stream capture, cudaStreamPerThread, more than kDynStreams streams.
unsigned long long* dyn_slot(DeviceInfo* di, cudaStream_t st, unsigned int* tag)
{
    if (st == cudaStreamPerThread) return nullptr;
    if (st == nullptr) st = cudaStreamLegacy;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (cap != cudaStreamCaptureStatusNone) return nullptr;
    uint32_t t = di->dyn_tag.fetch_add(1, std::memory_order_relaxed) + 1u;
    if (t == 0u) t = di->dyn_tag.fetch_add(1, std::memory_order_relaxed) + 1u;  // 0 means "free"
    *tag = t;
    std::lock_guard<std::mutex> lk(di->dyn_mu);
    auto it = di->dyn_map.find(st);
    if (it == di->dyn_map.end()) {
        if ((int)di->dyn_map.size() >= kDynStreams) return nullptr;
        it = di->dyn_map.emplace(st, DynStream{(int)di->dyn_map.size(), 0}).first;
    }
    const uint64_t k = it->second.count++;
    return reinterpret_cast<unsigned long long*>(di->dyn + ((size_t)it->second.base * 2 + (size_t)(k & 1u)) * kDynSlotBytes);
}

#endif  // FSB_ALL_VARIANTS

// Enqueue one kernel on the current device.
// overlap: launch with the programmatic-stream-serialization attribute, so that this kernel's
// CTAs may start (and read d_array) before the previous kernel of `st` has finished; everything
// it publishes still waits for that kernel (flagstat_kernels.cuh, "Overlapped steps").
std::atomic<int> g_pdl_unsupported{0};

int launch(int mode, const uint16_t* d_array, uint64_t n, uint64_t* d_out, cudaStream_t st,
           const XchgArgs* xa = nullptr, bool overlap = false)
{
    if ((reinterpret_cast<uintptr_t>(d_array) & 1u) != 0) return FLAGSTAT_CUDA_EINVAL;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    DeviceInfo* di = nullptr;
    int rc = device_info(dev, &di);
    if (rc) return rc;
    int variant = g_variant.load();
    if (variant == -1) {  // first launch: env FLAGSTAT_CUDA_VARIANT picks the default (A/B runs)
        variant = 0;
        if (const char* e = std::getenv("FLAGSTAT_CUDA_VARIANT")) variant = std::atoi(e);
        g_variant.store(variant);
    }
    if (!variant_built(variant)) variant = 0;
    const bool sam = mode == kSamtools;
    const KernelCfg& k = sam ? kSamtoolsKernel : kKernels[variant];
    KernelFn fn = sam ? k.fn[0] : k.fn[mode];

    const uint64_t addr = reinterpret_cast<uintptr_t>(d_array);
    uint64_t head = ((16u - (addr & 15u)) & 15u) >> 1;
    if (head > n) head = n;
    const uint64_t nvec = (n - head) >> 3;
    const uint64_t nb = nvec / kVecPerBatch;
    int per_sm = g_ctas_per_sm.load();
    const bool default_grid = per_sm <= 0;
    if (per_sm <= 0) per_sm = sam ? di->occ_samtools : di->occ[variant][mode];
    uint64_t grid = (uint64_t)per_sm * (uint64_t)di->sms;
    if (grid > nb + 1) grid = nb + 1;

    XchgArgs args;
    if (xa) args = *xa;
    else std::memset(&args, 0, sizeof(args));
    args.nb_q = nb / grid;
    args.nb_r = (unsigned)(nb % grid);
    args.pdl = 0;
    args.dyn = nullptr;
    // Columns long enough to keep every resident warp busy for several chunks take the dynamically
    // scheduled twin of the default kernel (same arithmetic, work claimed per warp from a counter).
#ifdef FSB_ALL_VARIANTS
    if ((variant == 0 || sam) && default_grid && dyn_min_chunks() >= 0) {
        const int cg = dyn_cg();
        const uint64_t chunk_vec = (uint64_t)cg * kDynWarpVecPerGroup;
        const uint64_t full = (uint64_t)di->occ_dyn[mode] * (uint64_t)di->sms;
        const uint64_t chunks = nvec / chunk_vec;
        const uint64_t need = dyn_min_chunks() > 0 ? (uint64_t)dyn_min_chunks() : 6ull * full * kWarps;
        if (chunks >= need && chunks < 0xFFFFFFF0ull) {
            if (unsigned long long* slot = dyn_slot(di, st, &args.dyn_tag)) {
                args.dyn = slot;
                fn = kDynKernels[cg - 1][mode];
                grid = full;
            }
        }
    }
#else
    (void)default_grid;
#endif
    if (overlap && !g_pdl_unsupported.load(std::memory_order_relaxed)) {
        args.pdl = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(k.threads);
        cfg.dynamicSmemBytes = k.smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, fn, d_array, (uint64_t)n,
                                                 reinterpret_cast<unsigned long long*>(d_out), args);
        if (e == cudaSuccess) {
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return 0;
        }
        cudaGetLastError();  // e.g. a stream kind that cannot take the attribute: plain launch from now on
        g_pdl_unsupported.store(1, std::memory_order_relaxed);
        args.pdl = 0;
    }
    fn<<<dim3((unsigned)grid), dim3(k.threads), k.smem, st>>>(
        d_array, n, reinterpret_cast<unsigned long long*>(d_out), args);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// per-call working set for synchronous entry points (pooled per device)
// ---------------------------------------------------------------------------
struct Lane {
    int dev = -1;
    cudaStream_t copy = nullptr, comp = nullptr;
    uint64_t* d_flags = nullptr;  // 32 x u64
    uint64_t* h_flags = nullptr;  // pinned, mapped
    unsigned long long* xbuf = nullptr;  // world-1 exchange buffer: lets one launch OVERWRITE its output
    uint16_t* stage[kStages] = {nullptr, nullptr, nullptr};
    cudaEvent_t copied[kStages] = {nullptr, nullptr, nullptr};
    cudaEvent_t consumed[kStages] = {nullptr, nullptr, nullptr};
    bool staged = false;
};

std::mutex g_pool_mu;
std::vector<Lane*> g_pool[kMaxDevices];

void lane_destroy(Lane* l)
{
    if (!l) return;
    for (int i = 0; i < kStages; ++i) {
        if (l->stage[i]) cudaFree(l->stage[i]);
        if (l->copied[i]) cudaEventDestroy(l->copied[i]);
        if (l->consumed[i]) cudaEventDestroy(l->consumed[i]);
    }
    if (l->xbuf) cudaFree(l->xbuf);
    if (l->d_flags) cudaFree(l->d_flags);
    if (l->h_flags) cudaFreeHost(l->h_flags);
    if (l->copy) cudaStreamDestroy(l->copy);
    if (l->comp) cudaStreamDestroy(l->comp);
    delete l;
}

int lane_create(int dev, Lane** out)
{
    Lane* l = new (std::nothrow) Lane();
    if (!l) return FLAGSTAT_CUDA_ENOMEM;
    l->dev = dev;
    auto build = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&l->copy, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&l->comp, cudaStreamNonBlocking));
        CK(cudaMalloc(&l->d_flags, 32 * sizeof(uint64_t)));
        // mapped: the one-launch path for short host blocks lets the kernel store the totals here
        CK(cudaHostAlloc(&l->h_flags, 32 * sizeof(uint64_t), cudaHostAllocMapped));
        CK(cudaMalloc(&l->xbuf, kXchgWords * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(l->xbuf, 0, kXchgWords * sizeof(unsigned long long), l->comp));
        CK(cudaStreamSynchronize(l->comp));
        return 0;
    };
    const int rc = build();
    if (rc) {  // nothing may leak from a half-built lane
        cudaGetLastError();
        lane_destroy(l);
        return rc;
    }
    *out = l;
    return 0;
}

int lane_ensure_staging(Lane* l)
{
    if (l->staged) return 0;
    for (int i = 0; i < kStages; ++i) {
        CK(cudaMalloc(&l->stage[i], kChunkBytes));
        CK(cudaEventCreateWithFlags(&l->copied[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&l->consumed[i], cudaEventDisableTiming));
    }
    l->staged = true;
    return 0;
}

int lane_acquire(int dev, Lane** out)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_pool[dev].empty()) {
            *out = g_pool[dev].back();
            g_pool[dev].pop_back();
            return 0;
        }
    }
    return lane_create(dev, out);
}

void lane_release(Lane* l)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool[l->dev].push_back(l);
}

// flagstat_blockfile.inl: T threads copy slices of a PAGEABLE host array into pinned slots
int run_pageable(int mode, const uint16_t* array, uint64_t len, uint64_t* totals);
// host arrays of at least this many bytes that are not page-locked take run_pageable()
uint64_t pageable_min_bytes()
{
    static const uint64_t v = [] {
        if (const char* e = std::getenv("FLAGSTAT_CUDA_PAGEABLE_MIN")) return (uint64_t)std::strtoull(e, nullptr, 10);
        // spawning the staging threads costs ~0.5 ms: below ~16 MB the plain cudaMemcpyAsync of
        // pageable memory is faster (4 M records: 1.11 ms threaded vs ~0.83 ms plain, r4c_dropin_time.jsonl)
        return (uint64_t)(16u << 20);
    }();
    return v;
}

// Synchronous run on the CURRENT device; array may be host or device memory.
// totals: 32 (flagstat) or 16 (pospopcnt) u64, overwritten.
int run_sync_impl(int mode, const uint16_t* array, uint64_t len, uint64_t* totals);
int run_sync(int mode, const uint16_t* array, uint64_t len, uint64_t* totals)
{
    return guarded([&] { return run_sync_impl(mode, array, len, totals); });
}

int run_sync_impl(int mode, const uint16_t* array, uint64_t len, uint64_t* totals)
{
    const int nout = mode == kPospopcnt ? 16 : 32;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (!array && len) return FLAGSTAT_CUDA_EINVAL;
    if ((reinterpret_cast<uintptr_t>(array) & 1u) != 0) return FLAGSTAT_CUDA_EINVAL;
    int dev = 0;
    CK(cudaGetDevice(&dev));

    bool on_device = false, pageable = false;
    struct Restore {  // device-resident input runs on the device that owns it
        int dev = -1;
        ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
    } restore;
    if (len) {
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, array);
        if (e == cudaSuccess) {
            on_device = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
            if (attr.type == cudaMemoryTypeDevice && attr.device != dev) {
                CK(cudaSetDevice(attr.device));
                restore.dev = dev;
                dev = attr.device;
            }
            pageable = attr.type == cudaMemoryTypeUnregistered;
        } else {
            cudaGetLastError();  // plain malloc memory on old drivers
            pageable = true;
        }
    }
    if (pageable && len * sizeof(uint16_t) >= pageable_min_bytes())
        return run_pageable(mode, array, len, totals);

    Lane* l = nullptr;
    int rc = lane_acquire(dev, &l);
    if (rc) return rc;
    struct Release {
        Lane* l;
        ~Release() { lane_release(l); }
    } rel{l};

    // Device-resident input: the caller's producer kernels were enqueued on
    // streams this library knows nothing about.  Run on the LEGACY default
    // stream, which is ordered after all prior work on blocking streams (the
    // semantics a synchronous C call on a device pointer is expected to have).
    // Host input is ready by definition, so it uses the lane's private streams.
    cudaStream_t comp = (on_device && len) ? cudaStreamLegacy : l->comp;
    const uint64_t chunk_rec = kChunkBytes / sizeof(uint16_t);
    if (on_device || len <= chunk_rec) {
        // ONE launch covers the call (device-resident input, or a host block of at most one
        // staging chunk -- the reference's 512,000-record blocks, benchmark/flagstats.cpp:304-329):
        // copy, then a launch in overwrite mode (world-1 exchange: the CTA that draws the last
        // ticket stores the totals) straight into the lane's MAPPED host counters.  No memset,
        // no device->host copy, no events: memcpy + launch + one synchronisation.
        const uint16_t* src = array;
        if (!on_device && len) {
            rc = lane_ensure_staging(l);
            if (rc) return rc;
            CK(cudaMemcpyAsync(l->stage[0], array, len * sizeof(uint16_t), cudaMemcpyHostToDevice, comp));
            src = l->stage[0];
        }
        uint64_t* d_host_flags = nullptr;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_host_flags), l->h_flags, 0));
        XchgArgs xa;
        std::memset(&xa, 0, sizeof(xa));
        xa.buf[0] = l->xbuf;
        xa.epoch = 1;
        xa.world = 1;
        rc = launch(mode, src, len, d_host_flags, comp, &xa);
        if (rc) return rc;
        CK(cudaStreamSynchronize(comp));
        std::memcpy(totals, l->h_flags, nout * sizeof(uint64_t));
        return 0;
    }
    CK(cudaMemsetAsync(l->d_flags, 0, 32 * sizeof(uint64_t), comp));
    {
        rc = lane_ensure_staging(l);
        if (rc) return rc;
        uint64_t off = 0;
        for (uint64_t c = 0; off < len; ++c) {
            const int s = (int)(c % kStages);
            const uint64_t n = (len - off < chunk_rec) ? (len - off) : chunk_rec;
            if (c >= (uint64_t)kStages) CK(cudaStreamWaitEvent(l->copy, l->consumed[s], 0));
            CK(cudaMemcpyAsync(l->stage[s], array + off, n * sizeof(uint16_t),
                               cudaMemcpyHostToDevice, l->copy));
            CK(cudaEventRecord(l->copied[s], l->copy));
            CK(cudaStreamWaitEvent(comp, l->copied[s], 0));
            rc = launch(mode, l->stage[s], n, l->d_flags, comp);
            if (rc) return rc;
            CK(cudaEventRecord(l->consumed[s], comp));
            off += n;
        }
    }
    CK(cudaMemcpyAsync(l->h_flags, l->d_flags, nout * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                       comp));
    CK(cudaStreamSynchronize(comp));
    std::memcpy(totals, l->h_flags, nout * sizeof(uint64_t));
    return 0;
}

uint32_t min_len_init()
{
    uint32_t v = g_min_len.load();
    if (v) return v;
    // Measured, not guessed (profiles/r4c_dropin_time.jsonl, r4e_dropin_time.jsonl: one synchronous
    // FLAGSTAT_cuda call per n records of PAGEABLE host memory -- what the reference's callers hand
    // over -- against FLAGSTAT_avx512 on one core of the same host): 524,288 records 129 vs 126 us
    // (a tie: the driver's single-threaded staging copy of pageable memory costs as much as the
    // AVX-512 kernel), 1,048,576 records 227 vs 253 us, 2,097,152 418 vs 506 us, and from 8 M records
    // on the threaded staging path pulls away (16.7 M: 1.85 vs 4.07 ms).  From PINNED memory the
    // call wins from 131,072 records (24.7 vs 31.7 us): callers that own pinned buffers lower the
    // threshold with FLAGSTAT_CUDA_MIN_LEN / FLAGSTAT_cuda_set_min_len.
    v = 1048576u;
    if (const char* e = std::getenv("FLAGSTAT_CUDA_MIN_LEN")) {
        const unsigned long long t = std::strtoull(e, nullptr, 10);
        if (t > 0 && t <= 0xFFFFFFFFull) v = (uint32_t)t;
    }
    g_min_len.store(v);
    return v;
}

}  // namespace

// ---------------------------------------------------------------------------
// streaming handle
// ---------------------------------------------------------------------------
struct FLAGSTAT_cuda_stream {
    int dev = 0;
    uint32_t block_records = 0;
    int n_groups = 0;     // ring depth, in groups
    int coalesce = 1;     // blocks per group = per DMA + kernel launch
    int mode = 0;         // FLAGSTAT_CUDA_STREAM_DMA / _ZEROCOPY
    uint16_t* h_base = nullptr;  // pinned: n_groups * coalesce blocks, contiguous
    uint16_t* d_base = nullptr;  // device twin (DMA mode only)
    std::vector<cudaStream_t> st;    // per group
    std::vector<cudaEvent_t> done;   // per group
    std::vector<char> busy;          // per group
    int cur_group = 0;
    int cur_fill = 0;            // blocks submitted into the current group
    uint64_t cur_records = 0;    // records submitted into the current group (contiguous)
    bool acquired = false;
    uint64_t* d_flags = nullptr;
    uint64_t* h_flags = nullptr;
};

namespace {

// Ship the current group: one DMA of its contiguous bytes (DMA mode) and one
// kernel launch over all of them, on the group's own stream.
int stream_flush_group(FLAGSTAT_cuda_stream* s)
{
    if (s->cur_fill == 0) return 0;
    const int g = s->cur_group;
    const size_t off = (size_t)g * s->coalesce * s->block_records;
    if (s->cur_records) {
        const uint16_t* src = s->h_base + off;
        if (s->mode == FLAGSTAT_CUDA_STREAM_DMA) {
            CK(cudaMemcpyAsync(s->d_base + off, src, s->cur_records * sizeof(uint16_t),
                               cudaMemcpyHostToDevice, s->st[g]));
            src = s->d_base + off;
        }
        // zero-copy: the kernel's cp.async loads pull the pinned block over PCIe themselves
        const int rc = launch(kFlagstat, src, s->cur_records, s->d_flags, s->st[g]);
        if (rc) return rc;
    }
    CK(cudaEventRecord(s->done[g], s->st[g]));
    s->busy[g] = 1;
    s->cur_group = (g + 1) % s->n_groups;
    s->cur_fill = 0;
    s->cur_records = 0;
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------
// fused counter exchange handle (one per rank)
// ---------------------------------------------------------------------------
struct FLAGSTAT_cuda_xchg {
    int dev = 0;
    int rank = 0;
    int world = 1;
    bool connected = false;
    bool ipc = false;
    unsigned long long* mine = nullptr;                 // this rank's buffer (cudaMalloc)
    unsigned long long* peer[fsb200::kMaxRanks] = {};  // every rank's buffer as mapped here
    uint64_t epoch = 0;
    uint64_t timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    bool overlap = false;  // FLAGSTAT_cuda_xchg_set_overlap
    // deferred collection: the epoch whose global counters the NEXT launch (or _collect) writes
    uint64_t pending_epoch = 0;
    uint64_t* pending_out = nullptr;
    int pending_accumulate = 0;
    int pending_nout = 0;
    unsigned long long* h_err = nullptr;  // mapped pinned word: the kernel stores the epoch of a timeout
};

extern "C" {

int FLAGSTAT_cuda_available(void) { return probe_devices() > 0 ? probe_devices() : 0; }

uint32_t FLAGSTAT_cuda_min_len(void) { return min_len_init(); }
void FLAGSTAT_cuda_set_min_len(uint32_t n) { g_min_len.store(n ? n : 1u); }

int FLAGSTAT_cuda_u64(const uint16_t* array, uint64_t len, uint64_t* flags)
{
    if (!flags) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[32];
    const int rc = run_sync(kFlagstat, array, len, t);
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) flags[i] += t[i];
    return 0;
}

int FLAGSTAT_cuda(const uint16_t* array, uint32_t len, uint32_t* flags)
{
    if (!flags) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[32];
    const int rc = run_sync(kFlagstat, array, len, t);
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) flags[i] += (uint32_t)t[i];  // wraps like the reference's ++
    return 0;
}

int FLAGSTAT_cuda_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, void* stream)
{
    if (!d_flags || (!d_array && len)) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    return launch(kFlagstat, d_array, len, d_flags, static_cast<cudaStream_t>(stream));
}

// ---- samtools mode: counters + exact n_pair_all (benchmark/flagstats.cpp:43-71) ----

int FLAGSTAT_cuda_samtools_u64(const uint16_t* array, uint64_t len, uint64_t* flags)
{
    if (!flags) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[32];
    const int rc = run_sync(kSamtools, array, len, t);
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) flags[i] += t[i];
    return 0;
}

int FLAGSTAT_cuda_samtools_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, void* stream)
{
    if (!d_flags || (!d_array && len)) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    return launch(kSamtools, d_array, len, d_flags, static_cast<cudaStream_t>(stream));
}

int FLAGSTAT_cuda_samtools_from_counters(const uint64_t* flags, FLAGSTAT_cuda_bam_flagstat* s)
{
    if (!flags || !s) return FLAGSTAT_CUDA_EINVAL;
    for (int w = 0; w < 2; ++w) {
        const uint64_t* f = flags + 16 * w;
        const uint64_t total = flags[w ? 25 : 9];
        s->n_reads[w] += (long long)total;
        s->n_mapped[w] += (long long)(total - f[2]);
        s->n_pair_all[w] += (long long)f[0];
        s->n_pair_map[w] += (long long)f[14];
        s->n_pair_good[w] += (long long)f[12];
        s->n_sgltn[w] += (long long)f[13];
        s->n_read1[w] += (long long)f[6];
        s->n_read2[w] += (long long)f[7];
        s->n_dup[w] += (long long)f[10];
        s->n_secondary[w] += (long long)f[8];
        s->n_supp[w] += (long long)f[11];
    }
    return 0;
}

int FLAGSTAT_cuda_samtools(const uint16_t* array, uint64_t len, FLAGSTAT_cuda_bam_flagstat* s)
{
    if (!s) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[32];
    const int rc = run_sync(kSamtools, array, len, t);
    if (rc) return rc;
    return FLAGSTAT_cuda_samtools_from_counters(t, s);
}

int FLAGSTAT_cuda_samtools_report(const FLAGSTAT_cuda_bam_flagstat* s, char* buf, size_t capacity)
{
    if (!s || (!buf && capacity)) return FLAGSTAT_CUDA_EINVAL;
    return guarded([&]() -> int {
    // percent(), benchmark/flagstats.cpp:73-78: float division, then * 100.0 in double
    char b0[32], b1[32];
    auto pct = [](char* b, long long n, long long total) -> const char* {
        if (total != 0) std::snprintf(b, 32, "%.2f%%", (float)n / total * 100.0);
        else std::strcpy(b, "N/A");
        return b;
    };
    std::string o;
    char line[160];
    auto put = [&](const char* fmt, long long a, long long b) {
        std::snprintf(line, sizeof line, fmt, a, b);
        o += line;
    };
    auto put_pct = [&](const char* fmt, const long long* v, const long long* tot) {
        std::snprintf(line, sizeof line, fmt, v[0], v[1], pct(b0, v[0], tot[0]), pct(b1, v[1], tot[1]));
        o += line;
    };
    // the lines of benchmark/flagstats.cpp:577-588 (the two diffchr lines are commented out there)
    put("%lld + %lld in total (QC-passed reads + QC-failed reads)\n", s->n_reads[0], s->n_reads[1]);
    put("%lld + %lld secondary\n", s->n_secondary[0], s->n_secondary[1]);
    put("%lld + %lld supplementary\n", s->n_supp[0], s->n_supp[1]);
    put("%lld + %lld duplicates\n", s->n_dup[0], s->n_dup[1]);
    put_pct("%lld + %lld mapped (%s : %s)\n", s->n_mapped, s->n_reads);
    put("%lld + %lld paired in sequencing\n", s->n_pair_all[0], s->n_pair_all[1]);
    put("%lld + %lld read1\n", s->n_read1[0], s->n_read1[1]);
    put("%lld + %lld read2\n", s->n_read2[0], s->n_read2[1]);
    put_pct("%lld + %lld properly paired (%s : %s)\n", s->n_pair_good, s->n_pair_all);
    put("%lld + %lld with itself and mate mapped\n", s->n_pair_map[0], s->n_pair_map[1]);
    put_pct("%lld + %lld singletons (%s : %s)\n", s->n_sgltn, s->n_pair_all);
    if (capacity) {
        const size_t n = o.size() < capacity - 1 ? o.size() : capacity - 1;
        std::memcpy(buf, o.data(), n);
        buf[n] = '\0';
    }
    return (int)o.size();
    });
}

int FLAGSTAT_cuda_device_overlapped(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, void* stream)
{
    if (!d_flags || (!d_array && len)) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    return launch(kFlagstat, d_array, len, d_flags, static_cast<cudaStream_t>(stream), nullptr, true);
}

int POSPOPCNT_cuda_u16_u64(const uint16_t* data, uint64_t len, uint64_t* out)
{
    if (!out) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[16];
    const int rc = run_sync(kPospopcnt, data, len, t);
    if (rc) return rc;
    for (int i = 0; i < 16; ++i) out[i] = t[i];  // zero-then-count, libalgebra.h:3498
    return 0;
}

int POSPOPCNT_cuda_u16(const uint16_t* data, size_t len, uint32_t* out)
{
    if (!out) return FLAGSTAT_CUDA_EINVAL;
    uint64_t t[16];
    const int rc = run_sync(kPospopcnt, data, len, t);
    if (rc) return rc;
    for (int i = 0; i < 16; ++i) out[i] = (uint32_t)t[i];
    return 0;
}

int POSPOPCNT_cuda_device(const uint16_t* d_data, uint64_t len, uint64_t* d_out, void* stream)
{
    if (!d_out || (!d_data && len)) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    return launch(kPospopcnt, d_data, len, d_out, static_cast<cudaStream_t>(stream));
}

// ---- streaming -------------------------------------------------------------

int FLAGSTAT_cuda_stream_open_ex(FLAGSTAT_cuda_stream** out, int device, uint32_t block_records,
                                 int n_slots, int mode, int coalesce)
{
    if (!out || block_records == 0) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (device < 0 || device >= g_ndev) return FLAGSTAT_CUDA_EINVAL;
    if (n_slots == 0) n_slots = 4;
    if (coalesce == 0) coalesce = 8;
    if (n_slots < 2 || n_slots > 64 || coalesce < 1 || coalesce > 256) return FLAGSTAT_CUDA_EINVAL;
    if (mode != FLAGSTAT_CUDA_STREAM_DMA && mode != FLAGSTAT_CUDA_STREAM_ZEROCOPY)
        return FLAGSTAT_CUDA_EINVAL;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return (int)guard.err;
    FLAGSTAT_cuda_stream* s = new (std::nothrow) FLAGSTAT_cuda_stream();
    if (!s) return FLAGSTAT_CUDA_ENOMEM;
    s->dev = device;
    s->block_records = block_records;
    s->n_groups = n_slots;
    s->coalesce = coalesce;
    s->mode = mode;
    s->st.assign(n_slots, nullptr);
    s->done.assign(n_slots, nullptr);
    s->busy.assign(n_slots, 0);
    const size_t bytes = (size_t)n_slots * coalesce * block_records * sizeof(uint16_t);
    // a failure half-way must not leak the pinned ring: _close() copes with a partly built handle
    auto build = [&]() -> int {
        CK(cudaHostAlloc(&s->h_base, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        if (mode == FLAGSTAT_CUDA_STREAM_DMA) CK(cudaMalloc(&s->d_base, bytes));
        for (int i = 0; i < n_slots; ++i) {
            CK(cudaStreamCreateWithFlags(&s->st[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming));
        }
        CK(cudaMalloc(&s->d_flags, 32 * sizeof(uint64_t)));
        CK(cudaMallocHost(&s->h_flags, 32 * sizeof(uint64_t)));
        CK(cudaMemset(s->d_flags, 0, 32 * sizeof(uint64_t)));
        return 0;
    };
    const int rc = build();
    if (rc) {
        cudaGetLastError();
        FLAGSTAT_cuda_stream_close(s);
        return rc;
    }
    *out = s;
    return 0;
}

int FLAGSTAT_cuda_stream_open(FLAGSTAT_cuda_stream** out, int device, uint32_t block_records,
                              int n_slots)
{
    int mode = FLAGSTAT_CUDA_STREAM_DMA, coalesce = 0;
    if (const char* e = std::getenv("FLAGSTAT_CUDA_STREAM_MODE")) mode = std::atoi(e);
    if (const char* e = std::getenv("FLAGSTAT_CUDA_STREAM_COALESCE")) coalesce = std::atoi(e);
    return FLAGSTAT_cuda_stream_open_ex(out, device, block_records, n_slots, mode, coalesce);
}

uint16_t* FLAGSTAT_cuda_stream_acquire(FLAGSTAT_cuda_stream* s)
{
    if (!s || s->acquired) return nullptr;
    const int g = s->cur_group;
    if (s->cur_fill == 0 && s->busy[g]) {
        if (cudaEventSynchronize(s->done[g]) != cudaSuccess) return nullptr;
        s->busy[g] = 0;
    }
    s->acquired = true;
    return s->h_base + ((size_t)g * s->coalesce + s->cur_fill) * s->block_records;
}

int FLAGSTAT_cuda_stream_submit(FLAGSTAT_cuda_stream* s, uint32_t n_records)
{
    if (!s || !s->acquired) return FLAGSTAT_CUDA_ESTATE;
    if (n_records > s->block_records) return FLAGSTAT_CUDA_EINVAL;
    s->acquired = false;
    s->cur_records += n_records;
    s->cur_fill += 1;
    // a short block ends the contiguous run: ship the group right away
    if (s->cur_fill == s->coalesce || n_records < s->block_records) {
        DeviceGuard guard(s->dev);
        if (guard.err != cudaSuccess) return (int)guard.err;
        return stream_flush_group(s);
    }
    return 0;
}

int FLAGSTAT_cuda_stream_push(FLAGSTAT_cuda_stream* s, const uint16_t* block, uint32_t n_records)
{
    if (!s || (!block && n_records)) return FLAGSTAT_CUDA_EINVAL;
    if (n_records > s->block_records) return FLAGSTAT_CUDA_EINVAL;
    uint16_t* dst = FLAGSTAT_cuda_stream_acquire(s);
    if (!dst) return FLAGSTAT_CUDA_ESTATE;
    std::memcpy(dst, block, (size_t)n_records * sizeof(uint16_t));
    return FLAGSTAT_cuda_stream_submit(s, n_records);
}

int FLAGSTAT_cuda_stream_finish(FLAGSTAT_cuda_stream* s, uint64_t* flags)
{
    if (!s || !flags) return FLAGSTAT_CUDA_EINVAL;
    if (s->acquired) return FLAGSTAT_CUDA_ESTATE;
    DeviceGuard guard(s->dev);
    if (guard.err != cudaSuccess) return (int)guard.err;
    const int rc = stream_flush_group(s);
    if (rc) return rc;
    for (int i = 0; i < s->n_groups; ++i) {
        CK(cudaStreamSynchronize(s->st[i]));
        s->busy[i] = 0;
    }
    CK(cudaMemcpy(s->h_flags, s->d_flags, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CK(cudaMemset(s->d_flags, 0, 32 * sizeof(uint64_t)));
    for (int i = 0; i < 32; ++i) flags[i] += s->h_flags[i];
    return 0;
}

int FLAGSTAT_cuda_stream_selftime(FLAGSTAT_cuda_stream* s, uint32_t n_blocks, uint64_t* flags,
                                  double* seconds)
{
    if (!s || !flags || !seconds) return FLAGSTAT_CUDA_EINVAL;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < n_blocks; ++i) {
        if (!FLAGSTAT_cuda_stream_acquire(s)) return FLAGSTAT_CUDA_ESTATE;
        const int rc = FLAGSTAT_cuda_stream_submit(s, s->block_records);
        if (rc) return rc;
    }
    const int rc = FLAGSTAT_cuda_stream_finish(s, flags);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int FLAGSTAT_cuda_stream_close(FLAGSTAT_cuda_stream* s)
{
    if (!s) return FLAGSTAT_CUDA_EINVAL;
    DeviceGuard guard(s->dev);
    for (int i = 0; i < s->n_groups; ++i) {
        if (s->st[i]) cudaStreamSynchronize(s->st[i]);
        if (s->done[i]) cudaEventDestroy(s->done[i]);
        if (s->st[i]) cudaStreamDestroy(s->st[i]);
    }
    if (s->h_base) cudaFreeHost(s->h_base);
    if (s->d_base) cudaFree(s->d_base);
    if (s->d_flags) cudaFree(s->d_flags);
    if (s->h_flags) cudaFreeHost(s->h_flags);
    delete s;
    return 0;
}

// ---- several GPUs, one process ----------------------------------------------

int FLAGSTAT_cuda_multi_u64(const uint16_t* array, uint64_t len, uint64_t* flags, int n_devices)
{
    if (!flags || (!array && len)) return FLAGSTAT_CUDA_EINVAL;
    const int have = probe_devices();
    if (have <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (n_devices <= 0 || n_devices > have) n_devices = have;
    return guarded([&]() -> int {
    std::vector<std::thread> th;
    std::vector<int> rcs(n_devices, 0);
    std::vector<uint64_t> part((size_t)n_devices * 32, 0);
    for (int g = 0; g < n_devices; ++g) {
        const uint64_t lo = (g * (len / n_devices)) & ~7ull;
        const uint64_t hi = (g == n_devices - 1) ? len : (((g + 1) * (len / n_devices)) & ~7ull);
        th.emplace_back([=, &rcs, &part]() {
            cudaError_t e = cudaSetDevice(g);
            if (e != cudaSuccess) {
                rcs[g] = (int)e;
                return;
            }
            rcs[g] = run_sync(kFlagstat, array + lo, hi - lo, &part[(size_t)g * 32]);
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < n_devices; ++g)
        if (rcs[g]) return rcs[g];
    for (int g = 0; g < n_devices; ++g)
        for (int i = 0; i < 32; ++i) flags[i] += part[(size_t)g * 32 + i];
    return 0;
    });
}

// ---- fused count + counter exchange over peer memory ------------------------------

int FLAGSTAT_cuda_xchg_create(FLAGSTAT_cuda_xchg** out, int rank, int world, void* handle_out)
{
    if (!out || world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
        return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    FLAGSTAT_cuda_xchg* x = new (std::nothrow) FLAGSTAT_cuda_xchg();
    if (!x) return FLAGSTAT_CUDA_ENOMEM;
    x->rank = rank;
    x->world = world;
    auto build = [&]() -> int {
        CK(cudaGetDevice(&x->dev));
        CK(cudaMalloc(&x->mine, kXchgWords * sizeof(unsigned long long)));
        CK(cudaMemset(x->mine, 0, kXchgWords * sizeof(unsigned long long)));
        CK(cudaHostAlloc(&x->h_err, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
        *x->h_err = 0ull;
        CK(cudaDeviceSynchronize());
        x->peer[rank] = x->mine;
        x->connected = world == 1;
        if (handle_out) {
            static_assert(sizeof(cudaIpcMemHandle_t) == FLAGSTAT_CUDA_XCHG_HANDLE_BYTES, "IPC handle size");
            cudaIpcMemHandle_t h;
            std::memset(&h, 0, sizeof(h));
            if (world > 1) CK(cudaIpcGetMemHandle(&h, x->mine));
            std::memcpy(handle_out, &h, sizeof(h));
        }
        return 0;
    };
    const int rc = build();
    if (rc) {  // nothing may leak from a half-built handle
        cudaGetLastError();
        if (x->mine) cudaFree(x->mine);
        if (x->h_err) cudaFreeHost(x->h_err);
        delete x;
        return rc;
    }
    *out = x;
    return 0;
}

int FLAGSTAT_cuda_xchg_connect(FLAGSTAT_cuda_xchg* x, const void* all_handles)
{
    if (!x || !all_handles) return FLAGSTAT_CUDA_EINVAL;
    if (x->connected) return 0;
    DeviceGuard guard(x->dev);
    if (guard.err != cudaSuccess) return (int)guard.err;
    const char* hs = static_cast<const char*>(all_handles);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hs + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {  // unmap what was mapped so far; the handle stays unconnected
            cudaGetLastError();
            for (int q = 0; q < r; ++q)
                if (q != x->rank && x->peer[q]) {
                    cudaIpcCloseMemHandle(x->peer[q]);
                    x->peer[q] = nullptr;
                }
            return (int)e;
        }
        x->peer[r] = static_cast<unsigned long long*>(p);
    }
    x->ipc = true;
    x->connected = true;
    return 0;
}

int FLAGSTAT_cuda_xchg_connect_local(FLAGSTAT_cuda_xchg** xs, int world)
{
    if (!xs || world < 1 || world > kMaxRanks) return FLAGSTAT_CUDA_EINVAL;
    for (int r = 0; r < world; ++r)
        if (!xs[r] || xs[r]->world != world || xs[r]->rank != r) return FLAGSTAT_CUDA_EINVAL;
    struct Restore {  // every return path puts the caller's current device back
        int dev = -1;
        ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
    } restore;
    CK(cudaGetDevice(&restore.dev));
    for (int a = 0; a < world; ++a) {
        CK(cudaSetDevice(xs[a]->dev));
        for (int b = 0; b < world; ++b) {
            if (xs[b]->dev != xs[a]->dev) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, xs[a]->dev, xs[b]->dev));
                if (!can) return FLAGSTAT_CUDA_ENODEV;
                cudaError_t e = cudaDeviceEnablePeerAccess(xs[b]->dev, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return (int)e;
            }
            xs[a]->peer[b] = xs[b]->mine;  // UVA: the same pointer is valid on every device
        }
        xs[a]->connected = true;
    }
    return 0;
}

static int xchg_fill(FLAGSTAT_cuda_xchg* x, XchgArgs& xa)
{
    if (!x->connected) return FLAGSTAT_CUDA_ESTATE;
    // an exchange that timed out stays failed: the epoch-parity argument needs every rank to have
    // completed every epoch.  All ranks destroy and recreate their handles (FLAGSTAT_cuda_xchg_status
    // tells a time-out from this state).
    if (*reinterpret_cast<volatile unsigned long long*>(x->h_err) != 0ull) return FLAGSTAT_CUDA_ESTATE;
    int cur = -1;
    CK(cudaGetDevice(&cur));
    if (cur != x->dev) return FLAGSTAT_CUDA_EINVAL;  // the caller launches on the handle's device
    std::memset(&xa, 0, sizeof(xa));
    for (int r = 0; r < x->world; ++r) xa.buf[r] = x->peer[r];
    xa.timeout_ns = x->timeout_ns;
    xa.rank = x->rank;
    xa.world = x->world;
    xa.prev_epoch = x->pending_epoch;
    xa.prev_out = reinterpret_cast<unsigned long long*>(x->pending_out);
    xa.prev_accumulate = x->pending_accumulate;
    xa.prev_nout = x->pending_nout;
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&xa.host_err), x->h_err, 0));
    return 0;
}

static int xchg_launch(FLAGSTAT_cuda_xchg* x, int mode, const uint16_t* d_array, uint64_t len,
                       uint64_t* d_out, int accumulate, void* stream, bool deferred = false)
{
    if (!x || !d_out || (!d_array && len)) return FLAGSTAT_CUDA_EINVAL;
    XchgArgs xa;
    int rc = xchg_fill(x, xa);
    if (rc) return rc;
    xa.epoch = x->epoch + 1;
    xa.accumulate = accumulate ? 1 : 0;
    xa.deferred = (deferred && x->world > 1) ? 1 : 0;  // one rank: nothing to wait for, written at once
    rc = launch(mode, d_array, len, d_out, static_cast<cudaStream_t>(stream), &xa, x->overlap);
    if (rc) return rc;  // a launch that never happened must not consume an epoch
    ++x->epoch;
    x->pending_epoch = 0;  // whatever was pending rides in this launch
    if (xa.deferred) {
        x->pending_epoch = x->epoch;
        x->pending_out = d_out;
        x->pending_accumulate = xa.accumulate;
        x->pending_nout = mode == kPospopcnt ? 16 : 32;
    }
    return 0;
}

int FLAGSTAT_cuda_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array, uint64_t len,
                                   uint64_t* d_flags, int accumulate, void* stream)
{
    return xchg_launch(x, kFlagstat, d_array, len, d_flags, accumulate, stream);
}

int FLAGSTAT_cuda_samtools_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array, uint64_t len,
                                            uint64_t* d_flags, int accumulate, void* stream)
{
    return xchg_launch(x, kSamtools, d_array, len, d_flags, accumulate, stream);
}

int POSPOPCNT_cuda_device_allreduce(FLAGSTAT_cuda_xchg* x, const uint16_t* d_data, uint64_t len,
                                    uint64_t* d_out, int accumulate, void* stream)
{
    return xchg_launch(x, kPospopcnt, d_data, len, d_out, accumulate, stream);
}

int FLAGSTAT_cuda_device_allreduce_deferred(FLAGSTAT_cuda_xchg* x, const uint16_t* d_array, uint64_t len,
                                            uint64_t* d_flags, int accumulate, void* stream)
{
    return xchg_launch(x, kFlagstat, d_array, len, d_flags, accumulate, stream, true);
}

int FLAGSTAT_cuda_xchg_collect(FLAGSTAT_cuda_xchg* x, void* stream)
{
    if (!x) return FLAGSTAT_CUDA_EINVAL;
    if (x->pending_epoch == 0) return 0;
    XchgArgs xa;
    const int rc = xchg_fill(x, xa);
    if (rc) return rc;
    xchg_collect_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(xa);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    x->pending_epoch = 0;
    return 0;
}

int FLAGSTAT_cuda_xchg_set_overlap(FLAGSTAT_cuda_xchg* x, int on)
{
    if (!x) return FLAGSTAT_CUDA_EINVAL;
    const int prev = x->overlap ? 1 : 0;
    x->overlap = on != 0;
    return prev;
}

int FLAGSTAT_cuda_xchg_set_timeout_ms(FLAGSTAT_cuda_xchg* x, uint32_t ms)
{
    if (!x) return FLAGSTAT_CUDA_EINVAL;
    x->timeout_ns = (uint64_t)ms * 1000000ull;
    return 0;
}

int FLAGSTAT_cuda_xchg_status(FLAGSTAT_cuda_xchg* x)
{
    if (!x) return FLAGSTAT_CUDA_EINVAL;
    DeviceGuard guard(x->dev);
    if (guard.err != cudaSuccess) return (int)guard.err;
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, x->mine + kXchgErr, sizeof(err), cudaMemcpyDeviceToHost));
    return err ? FLAGSTAT_CUDA_ETIMEOUT : 0;
}

int FLAGSTAT_cuda_xchg_destroy(FLAGSTAT_cuda_xchg* x)
{
    if (!x) return FLAGSTAT_CUDA_EINVAL;
    DeviceGuard guard(x->dev);
    cudaDeviceSynchronize();
    if (x->ipc)
        for (int r = 0; r < x->world; ++r)
            if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->mine) cudaFree(x->mine);
    if (x->h_err) cudaFreeHost(x->h_err);
    delete x;
    return 0;
}

// ---- diagnostics / support ----------------------------------------------------

const char* FLAGSTAT_cuda_strerror(int code)
{
    switch (code) {
        case 0: return "success";
        case FLAGSTAT_CUDA_ENODEV: return "no usable CUDA device";
        case FLAGSTAT_CUDA_EINVAL: return "invalid argument";
        case FLAGSTAT_CUDA_ENOMEM: return "host allocation failed";
        case FLAGSTAT_CUDA_ESTATE: return "handle used out of order";
        case FLAGSTAT_CUDA_ETIMEOUT: return "timed out waiting for a peer GPU's counters";
        case FLAGSTAT_CUDA_EFORMAT: return "malformed block container / LZ4 block";
        case FLAGSTAT_CUDA_EIO: return "cannot open or read the file";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

const char* FLAGSTAT_cuda_version(void) { return "libflagstats_cuda 0.1 (sm_100a)"; }

uint64_t FLAGSTAT_cuda_launch_count(void) { return g_launches.load(); }

int FLAGSTAT_cuda_set_variant(int v)
{
    if (!variant_built(v)) return FLAGSTAT_CUDA_EINVAL;  // not compiled into this build (see kKernels)
    const int prev = g_variant.exchange(v);
    return prev < 0 ? 0 : prev;
}
int FLAGSTAT_cuda_set_ctas_per_sm(int n) { return g_ctas_per_sm.exchange(n); }

int FLAGSTAT_cuda_set_dynamic(long long min_chunks, int groups_per_chunk)
{
#ifndef FSB_ALL_VARIANTS
    (void)groups_per_chunk;
    return min_chunks == -1 ? 0 : FLAGSTAT_CUDA_EINVAL;  // the product library has the static split only
#else
    if (min_chunks < -1 || (groups_per_chunk != 0 && groups_per_chunk != 1 && groups_per_chunk != 2))
        return FLAGSTAT_CUDA_EINVAL;
    g_dyn_min_chunks.store(min_chunks);
    if (groups_per_chunk) g_dyn_cg.store(groups_per_chunk);
    return 0;
#endif
}

const char* FLAGSTAT_cuda_kernel_name(int mode)
{
    // the instantiation launch() picks for the selected variant, spelled as ncu prints it
    static const char* const kNames[kNumVariants][2] = {
        {"fsb200::flagstat_kernel_group<0, 3, 2>", "fsb200::flagstat_kernel_group<1, 0, 2>"},
        {"fsb200::flagstat_kernel_group<0, 1, 2>", "fsb200::flagstat_kernel_group<1, 0, 2>"},
        {"fsb200::flagstat_kernel<0, 0>", "fsb200::flagstat_kernel<1, 0>"},
        {"fsb200::flagstat_kernel_tma<0, 0, 4, 2>", "fsb200::flagstat_kernel_tma<1, 0, 4, 2>"},
        {"fsb200::flagstat_kernel_tma<0, 0, 6, 2>", "fsb200::flagstat_kernel_tma<1, 0, 6, 2>"},
        {"fsb200::flagstat_kernel_ring<0, 0, 4, 2>", "fsb200::flagstat_kernel_ring<1, 0, 4, 2>"},
        {"fsb200::flagstat_kernel_ring<0, 0, 2, 2>", "fsb200::flagstat_kernel_ring<1, 0, 2, 2>"},
        {"fsb200::flagstat_kernel_ring<0, 2, 4, 2>", "fsb200::flagstat_kernel_ring<1, 0, 4, 2>"},
        {"fsb200::flagstat_kernel_group<0, 0, 2>", "fsb200::flagstat_kernel_group<1, 0, 2>"},
    };
    static_assert(kFlagstat == 0 && kPospopcnt == 1 && kSamtools == 2, "names above spell the MODE template argument");
    int v = g_variant.load();
    if (v == -1) {
        v = 0;
        if (const char* e = std::getenv("FLAGSTAT_CUDA_VARIANT")) v = std::atoi(e);
    }
    if (!variant_built(v)) v = 0;
    // long columns (>= 6 chunks per resident warp) of the default variant run its dynamically
    // scheduled twin; that is the instantiation the bench workloads launch
#ifdef FSB_ALL_VARIANTS
    if ((v == 0 || mode == kSamtools) && dyn_min_chunks() >= 0) {
        static const char* const kDyn[2][3] = {
            {"fsb200::flagstat_kernel_dyn<0, 3, 2, 1>", "fsb200::flagstat_kernel_dyn<1, 0, 2, 1>",
             "fsb200::flagstat_kernel_dyn<2, 3, 2, 1>"},
            {"fsb200::flagstat_kernel_dyn<0, 3, 2, 2>", "fsb200::flagstat_kernel_dyn<1, 0, 2, 2>",
             "fsb200::flagstat_kernel_dyn<2, 3, 2, 2>"}};
        return kDyn[dyn_cg() - 1][mode == kSamtools ? 2 : mode == kPospopcnt ? 1 : 0];
    }
#endif
    if (mode == kSamtools) return "fsb200::flagstat_kernel_group<2, 3, 2>";
    return kNames[v][mode == kPospopcnt ? 1 : 0];
}

int FLAGSTAT_cuda_synth_uniform(uint16_t* d_out, uint64_t start, uint64_t n, uint64_t seed,
                                uint16_t mask, void* stream)
{
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256 > 148u * 64u ? 148u * 64u : (n + 255) / 256);
    synth_uniform_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, start, n,
                                                                               seed, mask);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

int FLAGSTAT_cuda_synth_hiseqx(uint16_t* d_out, uint64_t start, uint64_t n, uint64_t seed,
                               uint32_t qcfail_ppm, void* stream)
{
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (n == 0) return 0;
    const unsigned grid = (unsigned)((n + 255) / 256 > 148u * 64u ? 148u * 64u : (n + 255) / 256);
    synth_hiseqx_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, start, n, seed,
                                                                              qcfail_ppm);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

void* FLAGSTAT_cuda_malloc(size_t bytes)
{
    void* p = nullptr;
    if (probe_devices() <= 0) return nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void* FLAGSTAT_cuda_malloc_host(size_t bytes)
{
    void* p = nullptr;
    if (probe_devices() <= 0) return nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

int FLAGSTAT_cuda_free(void* p) { CK(cudaFree(p)); return 0; }
int FLAGSTAT_cuda_free_host(void* p) { CK(cudaFreeHost(p)); return 0; }
int FLAGSTAT_cuda_memcpy_h2d(void* d, const void* h, size_t bytes)
{
    CK(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int FLAGSTAT_cuda_memcpy_d2h(void* h, const void* d, size_t bytes)
{
    CK(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int FLAGSTAT_cuda_memset(void* d, int value, size_t bytes)
{
    CK(cudaMemset(d, value, bytes));
    return 0;
}
int FLAGSTAT_cuda_sync(void)
{
    CK(cudaDeviceSynchronize());
    return 0;
}

int FLAGSTAT_cuda_time_device_rot(const uint16_t* d_base, uint64_t len, uint64_t stride_records,
                                  uint32_t n_rot, uint64_t* d_flags, int iters, int mode,
                                  float* ms_per_launch)
{
    if (!d_flags || !ms_per_launch || iters <= 0 || n_rot == 0) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    const bool overlapped = (mode & 4) != 0;  // launches carry the programmatic-serialization attribute
    const int m = (mode & 3) == 2 ? kSamtools : (mode & 3) ? kPospopcnt : kFlagstat;
    TimerPair t;
    CK(cudaStreamCreateWithFlags(&t.st, cudaStreamNonBlocking));
    t.own_stream = true;
    CK(cudaEventCreate(&t.e0));
    CK(cudaEventCreate(&t.e1));
    int rc = 0;
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(t.e0, t.st));
    for (int i = 0; i < iters && rc == 0; ++i)
        rc = launch(m, d_base + (uint64_t)((uint32_t)i % n_rot) * stride_records, len, d_flags, t.st, nullptr,
                    overlapped);
    CK(cudaEventRecord(t.e1, t.st));
    CK(cudaStreamSynchronize(t.st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t.e0, t.e1));
    *ms_per_launch = ms / (float)iters;
    return rc;
}

int FLAGSTAT_cuda_time_device(const uint16_t* d_array, uint64_t len, uint64_t* d_flags, int iters,
                              int pospopcnt_mode, float* ms_per_launch)
{
    return FLAGSTAT_cuda_time_device_rot(d_array, len, 0, 1, d_flags, iters, pospopcnt_mode, ms_per_launch);
}

// Read-only HBM probe: the same bytes through LDG.128 with one XOR per 16 bytes and no
// other work (tools/hbm_read_probe.cu holds the sweep this configuration won: 8 loads in
// flight per thread, 4 CTAs per SM).  bench.py reports the flagstat kernel against it
// next to the copy-kernel peak of MEASURED_PEAKS.json.
int FLAGSTAT_cuda_read_probe(const void* d_bytes, uint64_t n_bytes, int iters, float* ms_per_launch)
{
    if (!d_bytes || !ms_per_launch || iters <= 0) return FLAGSTAT_CUDA_EINVAL;
    if ((reinterpret_cast<uintptr_t>(d_bytes) & 15u) != 0) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    DeviceInfo* di = nullptr;
    int rc = device_info(dev, &di);
    if (rc) return rc;
    Lane* l = nullptr;
    rc = lane_acquire(dev, &l);
    if (rc) return rc;
    struct Release {
        Lane* l;
        ~Release() { lane_release(l); }
    } rel{l};
    TimerPair t;
    CK(cudaEventCreate(&t.e0));
    CK(cudaEventCreate(&t.e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(t.e0, l->comp));
    for (int i = 0; i < iters; ++i)
        hbm_read_probe_kernel<<<di->sms * 4, kThreads, 0, l->comp>>>(
            static_cast<const uint4*>(d_bytes), n_bytes / 16u,
            reinterpret_cast<unsigned long long*>(l->d_flags));
    CK(cudaEventRecord(t.e1, l->comp));
    CK(cudaStreamSynchronize(l->comp));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t.e0, t.e1));
    *ms_per_launch = ms / (float)iters;
    return 0;
}

#ifdef FSB_LZ4_PROFILE
// probe build only (tools/lz4_phase_probe.py): clock64() sums per phase of the LZ4 group decoder
int FLAGSTAT_cuda_lz4_profile_fetch(unsigned long long* out16, int clear)
{
    if (!out16) return FLAGSTAT_CUDA_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out16, fsb200::g_lz4_prof, 16 * sizeof(unsigned long long)));
    if (clear) {
        unsigned long long z[16] = {0};
        CK(cudaMemcpyToSymbol(fsb200::g_lz4_prof, z, sizeof(z)));
    }
    return 0;
}
// the same for the CTA decoder (lz4_block_cta.cuh): thread 0 of every CTA
int FLAGSTAT_cuda_l4_profile_fetch(unsigned long long* out16, int clear)
{
    if (!out16) return FLAGSTAT_CUDA_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out16, fsb200::g_l4_prof, 16 * sizeof(unsigned long long)));
    if (clear) {
        unsigned long long z[16] = {0};
        CK(cudaMemcpyToSymbol(fsb200::g_l4_prof, z, sizeof(z)));
    }
    return 0;
}
#endif

#ifdef FSB_TIMELINE
// probe build only (tools/timeline_probe.py): the per-CTA time stamps of the last group-kernel launch
int FLAGSTAT_cuda_timeline_fetch(unsigned long long* out, int n_words)
{
    if (!out || n_words <= 0 || n_words > 8 * 2048) return FLAGSTAT_CUDA_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, fsb200::g_timeline, (size_t)n_words * sizeof(unsigned long long)));
    return 0;
}
int FLAGSTAT_cuda_timeline_clear(void)
{
    static unsigned long long zeros[8 * 2048];
    CK(cudaMemcpyToSymbol(fsb200::g_timeline, zeros, sizeof(zeros)));
    return 0;
}
#endif

}  // extern "C"

#include "flagstat_blockfile.inl"

// ---- FLAG ingest (benchmark/utility.cpp:29-32) ---------------------------------------------
extern "C" int FLAGSTAT_cuda_ingest_text(const char* text, uint64_t n_bytes, uint16_t* out,
                                         uint64_t out_capacity, uint64_t* n_records, uint64_t* flags)
{
    using namespace fsb200;
    if ((!text && n_bytes) || !n_records) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    *n_records = 0;
    if (n_bytes == 0) return 0;
    const uint64_t tiles = (n_bytes + kIngestTile - 1) / kIngestTile;
    if (tiles > 0x7FFFFFFFull) return FLAGSTAT_CUDA_EINVAL;
    // a last line without '\n' is still a line (std::getline); parse it on the host
    uint64_t tail_lo = n_bytes;
    bool has_tail = text[n_bytes - 1] != '\n';
    if (has_tail)
        while (tail_lo > 0 && text[tail_lo - 1] != '\n') --tail_lo;
    unsigned char* d_text = nullptr;
    unsigned long long *d_tiles = nullptr, *d_first = nullptr;
    uint16_t* d_out = nullptr;
    void* d_tmp = nullptr;
    uint64_t* d_flags = nullptr;
    int rc = 0;
    uint64_t lines = 0;
    do {
        if ((rc = (int)cudaMalloc(&d_text, n_bytes))) break;
        if ((rc = (int)cudaMalloc(&d_tiles, (tiles + 1) * sizeof(unsigned long long)))) break;
        if ((rc = (int)cudaMalloc(&d_first, (tiles + 1) * sizeof(unsigned long long)))) break;
        if ((rc = (int)cudaMemcpy(d_text, text, n_bytes, cudaMemcpyHostToDevice))) break;
        if ((rc = (int)cudaMemset(d_tiles + tiles, 0, sizeof(unsigned long long)))) break;
        ingest_count_kernel<<<(unsigned)tiles, kIngestThreads>>>(d_text, n_bytes, d_tiles);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        size_t tmp_bytes = 0;
        if ((rc = (int)cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_tiles, d_first, (int)(tiles + 1)))) break;
        if ((rc = (int)cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1))) break;
        if ((rc = (int)cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_tiles, d_first, (int)(tiles + 1)))) break;
        unsigned long long total = 0;
        if ((rc = (int)cudaMemcpy(&total, d_first + tiles, sizeof(total), cudaMemcpyDeviceToHost))) break;
        lines = total + (has_tail ? 1 : 0);
        if (out && lines > out_capacity) {
            rc = FLAGSTAT_CUDA_EINVAL;
            *n_records = lines;  // tell the caller how much room is needed
            break;
        }
        if ((rc = (int)cudaMalloc(&d_out, (lines ? lines : 1) * sizeof(uint16_t)))) break;
        if (total) {
            ingest_parse_kernel<<<(unsigned)tiles, kIngestThreads>>>(d_text, n_bytes, d_first, d_out);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if ((rc = (int)cudaGetLastError())) break;
        }
        if (has_tail) {
            // atoi stops at the first non-digit after optional white space and sign: 32 characters
            // are more than it can consume of a value that is then truncated to 16 bits
            char last[40];
            uint64_t ws = tail_lo;  // (leading white space of any length is skipped here, as atoi would)
            while (ws < n_bytes && (text[ws] == ' ' || (text[ws] >= '\t' && text[ws] <= '\r'))) ++ws;
            const uint64_t tl = n_bytes - ws < sizeof(last) - 1 ? n_bytes - ws : sizeof(last) - 1;
            std::memcpy(last, text + ws, tl);
            last[tl] = '\0';
            const uint16_t v = (uint16_t)std::atoi(last);
            if ((rc = (int)cudaMemcpy(d_out + total, &v, sizeof(v), cudaMemcpyHostToDevice))) break;
        }
        if (flags) {  // count straight from the device column: the text never becomes a host uint16 array
            if ((rc = (int)cudaMalloc(&d_flags, 32 * sizeof(uint64_t)))) break;
            if ((rc = (int)cudaMemset(d_flags, 0, 32 * sizeof(uint64_t)))) break;
            if ((rc = launch(kFlagstat, d_out, lines, d_flags, nullptr))) break;
            uint64_t t[32];
            if ((rc = (int)cudaMemcpy(t, d_flags, sizeof(t), cudaMemcpyDeviceToHost))) break;
            for (int i = 0; i < 32; ++i) flags[i] += t[i];
        }
        if (out && lines)
            if ((rc = (int)cudaMemcpy(out, d_out, lines * sizeof(uint16_t), cudaMemcpyDeviceToHost))) break;
        if ((rc = (int)cudaDeviceSynchronize())) break;
        *n_records = lines;
    } while (0);
    cudaFree(d_text);
    cudaFree(d_tiles);
    cudaFree(d_first);
    cudaFree(d_out);
    cudaFree(d_tmp);
    cudaFree(d_flags);
    return rc;
}
