// flagstat_blockfile.inl -- consumer of the reference's FLAG files (included by
// flagstat_capi.cu inside extern "C").
//
//   FLAGSTAT_CUDA_FILE_RAW  plain uint16 stream (".bin"), which the reference reads
//                           in 1,024,000-byte blocks (benchmark/flagstats.cpp:415-468)
//   FLAGSTAT_CUDA_FILE_LZ4  [int32 raw_size][int32 comp_size][LZ4 block] records
//                           (written :110-186, read :288-358)
//
// Raw files: T reader threads pread() groups of blocks straight into private pinned slots, one DMA +
// one kernel launch per slot (raw_ctx / consume_raw below).  LZ4 and Zstd containers are shipped
// COMPRESSED: the payloads of a batch are gathered into a pinned staging buffer by every CPU of the
// affinity mask, cross PCIe in one DMA, are expanded in HBM -- LZ4 blocks by one CTA per block
// (lz4_block_cta.cuh), Zstd frames by an entropy stage + the same copy phase (zstd_block.cuh) -- and the
// flagstat kernel reads the decoded records from there.  Three staging lanes: batch k is decoded while
// k + 1 crosses PCIe and k + 2 is gathered on the host.

namespace {

struct ByteSource {
    FILE* fp = nullptr;
    const unsigned char* mem = nullptr;
    uint64_t size = 0, pos = 0;
    size_t read(void* dst, size_t n)
    {
        if (fp) {
            const size_t got = std::fread(dst, 1, n, fp);
            pos += got;
            return got;
        }
        const uint64_t left = size - pos;
        const size_t got = (size_t)(n < left ? n : left);
        std::memcpy(dst, mem + pos, got);
        pos += got;
        return got;
    }
    bool at_end()
    {
        if (!fp) return pos >= size;
        const int c = std::fgetc(fp);
        if (c == EOF) return true;
        std::ungetc(c, fp);
        return false;
    }
};

constexpr uint32_t kRefBlockBytes = 1024000u;          // benchmark/flagstats.cpp:119
constexpr uint32_t kMaxRawBlock = 8u << 20;            // sanity bound on a header's raw_size
// Upper bounds of a batch: kBatchBlocks blocks / kBatchRawCap decoded bytes.  (The warp-per-block A/B
// decoders need thousands of blocks per launch to fill the GPU and take batches of this size; the CTA
// decoder and the Zstd stages get wave-sized batches, see consume_lz4.)
constexpr int kBatchBlocks = 2048;
constexpr size_t kBatchRawCap = (size_t)1 << 31;

bool file_debug()
{
    static const bool on = std::getenv("FLAGSTAT_CUDA_DEBUG") != nullptr;
    return on;
}

// Staging / reader threads.  Default: 3/4 of the CPUs this process may run on (sched_getaffinity,
// not the machine's core count: containers and taskset), at least 2, at most kRawMaxThreads --
// measured on a 16-vCPU B200 host: 6 threads 43 GB/s, 12 threads 45-50 GB/s of a 55 GB/s link
// (profiles/r4c_pageable.jsonl).  FLAGSTAT_CUDA_IO_THREADS overrides.
int io_threads()
{
    int t = 0;
    if (const char* e = std::getenv("FLAGSTAT_CUDA_IO_THREADS")) t = std::atoi(e);
    if (t <= 0) {
        int cpus = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
        if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
        t = cpus * 3 / 4;
        if (t < 2) t = 2;
    }
    if (t < 1) t = 1;
    if (t > 24) t = 24;
    return t;
}

// Threads for gathering compressed payloads into a pinned staging buffer: every CPU of the affinity mask
// (the calling thread waits for them anyway), at most 24.  FLAGSTAT_CUDA_IO_THREADS overrides.
int gather_threads()
{
    if (const char* e = std::getenv("FLAGSTAT_CUDA_IO_THREADS")) {
        const int t = std::atoi(e);
        if (t > 0) return t > 24 ? 24 : t;
    }
    int cpus = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
    if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
    if (cpus < 2) cpus = 2;
    return cpus > 24 ? 24 : cpus;
}

// Copy into a pinned staging slot with non-temporal stores: the slot is written once and then
// read by the DMA engine, never by this core, so the destination lines need not be read for
// ownership nor kept in cache -- a quarter less DRAM traffic than memcpy on the staged path
// (source read + slot write + DMA read instead of + the slot's read-for-ownership), which is
// what bounds it once enough threads are copying.  dst is 64-byte aligned (slot base + a
// multiple of 64 KiB).  FLAGSTAT_CUDA_STAGING_NT=0 switches back to memcpy (A/B).
bool staging_nt()
{
    const char* e = std::getenv("FLAGSTAT_CUDA_STAGING_NT");
    return !(e && std::atoi(e) == 0);
}

void copy_streaming(unsigned char* dst, const unsigned char* src, size_t n)
{
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0u) {
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
        }
        _mm_sfence();
    }
    if (i < n) std::memcpy(dst + i, src + i, n - i);
}

// FLAGSTAT_CUDA_LZ4_VARIANT: 2 = one CTA per block, parse / copy phases (lz4_block_cta.cuh, default),
// 1 = one warp per block, 32 sequences per warp step (lz4_block_group.cuh), 0 = one warp per block,
// one sequence per step (lz4_block.cuh); the warp-per-block decoders are kept for A/B (profiles/)
std::atomic<int> g_lz4_variant{-1};
int lz4_variant()
{
    int v = g_lz4_variant.load();
    if (v < 0) {
        const char* e = std::getenv("FLAGSTAT_CUDA_LZ4_VARIANT");
        v = e ? std::atoi(e) : 2;
        if (v < 0 || v > 2) v = 2;
        g_lz4_variant.store(v);
    }
    return v;
}

// the CTA decoder's descriptor scratch: one region per CTA of the grid, sized for the largest block
struct Lz4Scratch {
    unsigned char* d = nullptr;
    size_t cap = 0;
};

unsigned lz4_cta_grid(uint32_t n_blocks, int sms)
{
    const unsigned full = 2u * (unsigned)sms;  // two CTAs (kL4Smem = 112,896 bytes of shared memory each) per SM
    return n_blocks < full ? n_blocks : full;
}

int lz4_scratch_reserve(Lz4Scratch& s, uint32_t n_blocks, uint32_t max_comp, uint32_t max_raw, int sms)
{
    const size_t need = (size_t)lz4_cta_grid(n_blocks, sms) * l4_scratch_bytes(max_comp, max_raw);
    if (need <= s.cap) return 0;
    if (s.d) cudaFree(s.d);
    s.d = nullptr;
    s.cap = 0;
    CK(cudaMalloc(&s.d, need + need / 4));
    s.cap = need + need / 4;
    return 0;
}

// max_comp / max_raw: the largest block of the batch (sizes the CTA decoder's scratch)
int lz4_launch(const unsigned char* d_comp, unsigned char* d_raw, const Lz4BlockDesc* d_desc, int* d_status,
               uint32_t n_blocks, cudaStream_t st, Lz4Scratch& scratch, uint32_t max_comp, uint32_t max_raw)
{
    // both decoders need more than 48 KiB of dynamic shared memory: that opt-in is a PER-DEVICE
    // function attribute, set for every device in device_info()
    int dev = 0;
    CK(cudaGetDevice(&dev));
    DeviceInfo* di = nullptr;
    const int rc = device_info(dev, &di);
    if (rc) return rc;
    if (lz4_variant() == 2) {
        const int src = lz4_scratch_reserve(scratch, n_blocks, max_comp, max_raw, di->sms);
        if (src) return src;
        const size_t stride = l4_scratch_bytes(max_comp, max_raw);
        const uint32_t desc_cap = max_comp / 3u + 64u;
        lz4_decode_cta_kernel<<<lz4_cta_grid(n_blocks, di->sms), kL4Threads, kL4Smem, st>>>(
            d_comp, d_raw, d_desc, d_status, n_blocks, scratch.d, stride, desc_cap);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        CK(cudaGetLastError());
        return 0;
    }
    const unsigned grid = (n_blocks + kLz4WarpsPerCta - 1) / kLz4WarpsPerCta;
    if (lz4_variant() == 0)
        lz4_decode_kernel<<<grid, kLz4WarpsPerCta * 32, kLz4Smem, st>>>(d_comp, d_raw, d_desc, d_status, n_blocks);
    else
        lz4_decode_group_kernel<<<grid, kLz4WarpsPerCta * 32, kLz4GroupSmem, st>>>(d_comp, d_raw, d_desc,
                                                                                  d_status, n_blocks);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

// codec of the [int32 raw][int32 comp][payload] records
enum Codec { kCodecLz4 = 0, kCodecZstd = 1 };

// Zstd frames (zstd_frame.cuh): one frame per CTA slot, per-frame workspace in global memory
int zstd_launch(const unsigned char* d_comp, unsigned char* d_raw, const Lz4BlockDesc* d_desc, int* d_status,
                uint32_t n_blocks, zstd::Work* d_work, cudaStream_t st)
{
    zstd_decode_kernel<<<n_blocks, 32, 0, st>>>(d_comp, d_raw, d_desc, d_status, n_blocks, d_work);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

// Zstd frames, second version (default): entropy stage -> descriptors + literals, then the LZ4 copy phase
// (zstd_block.cuh).  FLAGSTAT_CUDA_ZSTD_VARIANT=0 runs the first version (one thread per frame) for A/B.
int zstd_variant()
{
    const char* e = std::getenv("FLAGSTAT_CUDA_ZSTD_VARIANT");
    return (e && std::atoi(e) == 0) ? 0 : 1;
}

struct Zstd2 {   // buffers of one batch; grow on demand
    ZstdFrameAux* h_aux = nullptr;  // pinned
    ZstdFrameAux* d_aux = nullptr;
    ZstdParsed* d_parsed = nullptr;
    unsigned char* d_lit = nullptr;
    L4Desc* d_descs = nullptr;
    uint32_t* d_tf = nullptr;
    size_t blk_cap = 0, lit_cap = 0, desc_cap = 0, tf_cap = 0;
};

void zstd2_free(Zstd2& z)
{
    if (z.h_aux) cudaFreeHost(z.h_aux);
    if (z.d_aux) cudaFree(z.d_aux);
    if (z.d_parsed) cudaFree(z.d_parsed);
    if (z.d_lit) cudaFree(z.d_lit);
    if (z.d_descs) cudaFree(z.d_descs);
    if (z.d_tf) cudaFree(z.d_tf);
    z = Zstd2();
}

// h_comp: the batch's compressed bytes on the HOST (the frames' headers say how many descriptors each frame
// can produce: zstd::count_descriptors, header arithmetic only), h_desc: its descriptors.
int zstd2_launch(Zstd2& z, const unsigned char* h_comp, const Lz4BlockDesc* h_desc, const unsigned char* d_comp,
                 unsigned char* d_raw, const Lz4BlockDesc* d_desc, int* d_status, uint32_t n, cudaStream_t st)
{
    int dev = 0;
    DeviceInfo* di = nullptr;
    CK(cudaGetDevice(&dev));
    const int irc = device_info(dev, &di);
    if (irc) return irc;
    if (n > z.blk_cap) {
        if (z.h_aux) cudaFreeHost(z.h_aux);
        if (z.d_aux) cudaFree(z.d_aux);
        if (z.d_parsed) cudaFree(z.d_parsed);
        z.h_aux = nullptr; z.d_aux = nullptr; z.d_parsed = nullptr;
        z.blk_cap = 0;
        const size_t cap = (size_t)n + n / 4 + 16;
        CK(cudaMallocHost(&z.h_aux, cap * sizeof(ZstdFrameAux)));
        CK(cudaMalloc(&z.d_aux, cap * sizeof(ZstdFrameAux)));
        CK(cudaMalloc(&z.d_parsed, cap * sizeof(ZstdParsed)));
        z.blk_cap = cap;
    }
    size_t descs = 0, lits = 0;
    uint32_t max_raw = 0;
    for (uint32_t b = 0; b < n; ++b) {
        const uint64_t cnt = zstd::count_descriptors(h_comp + h_desc[b].comp_off, h_desc[b].comp_size, h_desc[b].raw_size);
        const uint32_t lit_cap = h_desc[b].raw_size + zstd::kBlockMax + 64u;
        z.h_aux[b] = ZstdFrameAux{descs, lits, (uint32_t)cnt, lit_cap};
        descs += (size_t)cnt;
        lits += ((size_t)lit_cap + 15u) & ~(size_t)15u;
        if (h_desc[b].raw_size > max_raw) max_raw = h_desc[b].raw_size;
    }
    if (descs > z.desc_cap) {
        if (z.d_descs) cudaFree(z.d_descs);
        z.d_descs = nullptr;
        z.desc_cap = 0;
        const size_t cap = descs + descs / 4 + 1024;
        CK(cudaMalloc(&z.d_descs, cap * sizeof(L4Desc)));
        z.desc_cap = cap;
    }
    if (lits > z.lit_cap) {
        if (z.d_lit) cudaFree(z.d_lit);
        z.d_lit = nullptr;
        z.lit_cap = 0;
        const size_t cap = lits + lits / 8 + 4096;
        CK(cudaMalloc(&z.d_lit, cap));
        z.lit_cap = cap;
    }
    const unsigned grid = lz4_cta_grid(n, di->sms);
    const uint32_t tf_stride = (max_raw + 15u) / kL4Tile + 8u;
    if ((size_t)grid * tf_stride > z.tf_cap) {
        if (z.d_tf) cudaFree(z.d_tf);
        z.d_tf = nullptr;
        z.tf_cap = 0;
        const size_t cap = (size_t)2 * di->sms * tf_stride * 2;
        CK(cudaMalloc(&z.d_tf, cap * sizeof(uint32_t)));
        z.tf_cap = cap;
    }
    CK(cudaMemcpyAsync(z.d_aux, z.h_aux, n * sizeof(ZstdFrameAux), cudaMemcpyHostToDevice, st));
    zstd_parse_kernel<<<(n + kZstdFramesPerCta - 1) / kZstdFramesPerCta, 32 * kZstdFramesPerCta, kZstdParseSmem, st>>>(
        d_comp, d_desc, z.d_aux, n, z.d_lit, z.d_descs, z.d_parsed);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    zstd_copy_kernel<<<grid, kL4Threads, kL4Smem, st>>>(d_raw, d_desc, z.d_aux, z.d_parsed, d_status, n, z.d_lit,
                                                        z.d_descs, z.d_tf, tf_stride);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    return 0;
}

struct BlockInfo {
    uint64_t src_off;  // payload position in the file / memory image
    uint32_t comp, raw;
};

// staging of one batch; capacities grow on demand and the buffers are pooled across calls
struct Lz4Lane {
    unsigned char* h_comp = nullptr;   // pinned
    unsigned char* d_comp = nullptr;
    unsigned char* d_raw = nullptr;
    Lz4BlockDesc* h_desc = nullptr;    // pinned
    Lz4BlockDesc* d_desc = nullptr;
    int* h_status = nullptr;           // pinned
    int* d_status = nullptr;
    zstd::Work* d_work = nullptr;      // Zstd, first version only: one workspace per block of the batch
    Zstd2 z2;                          // Zstd (default): descriptors, literals, per-frame results
    Lz4Scratch scratch;                // LZ4 CTA decoder: descriptor scratch
    uint32_t max_comp = 0, max_raw = 0;  // largest block of the batch in flight
    size_t comp_cap = 0, raw_cap = 0;
    int blk_cap = 0, work_cap = 0;
    cudaStream_t st = nullptr;
    int n = 0;            // blocks in flight
    bool busy = false;
};

void lz4_lane_free(Lz4Lane& l)
{
    if (l.st) cudaStreamSynchronize(l.st);
    if (l.h_comp) cudaFreeHost(l.h_comp);
    if (l.d_comp) cudaFree(l.d_comp);
    if (l.d_raw) cudaFree(l.d_raw);
    if (l.h_desc) cudaFreeHost(l.h_desc);
    if (l.d_desc) cudaFree(l.d_desc);
    if (l.h_status) cudaFreeHost(l.h_status);
    if (l.d_status) cudaFree(l.d_status);
    if (l.d_work) cudaFree(l.d_work);
    zstd2_free(l.z2);
    if (l.scratch.d) cudaFree(l.scratch.d);
    if (l.st) cudaStreamDestroy(l.st);
    l = Lz4Lane();
}

int lz4_lane_reserve(Lz4Lane& l, size_t comp_bytes, size_t raw_bytes, int blocks, int codec)
{
    if (!l.st) CK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
    if (codec == kCodecZstd && zstd_variant() == 0 && blocks > l.work_cap) {
        if (l.d_work) cudaFree(l.d_work);
        l.d_work = nullptr;
        l.work_cap = 0;
        const int cap = blocks + blocks / 4 + 16;
        CK(cudaMalloc(&l.d_work, (size_t)cap * sizeof(zstd::Work)));
        l.work_cap = cap;
    }
    if (comp_bytes > l.comp_cap) {
        if (l.h_comp) cudaFreeHost(l.h_comp);
        if (l.d_comp) cudaFree(l.d_comp);
        l.h_comp = nullptr;
        l.d_comp = nullptr;
        l.comp_cap = 0;
        const size_t cap = comp_bytes + comp_bytes / 8 + 4096;
        CK(cudaMallocHost(&l.h_comp, cap));
        CK(cudaMalloc(&l.d_comp, cap));
        l.comp_cap = cap;
    }
    if (raw_bytes > l.raw_cap) {
        if (l.d_raw) cudaFree(l.d_raw);
        l.d_raw = nullptr;
        l.raw_cap = 0;
        const size_t cap = raw_bytes + raw_bytes / 8 + 4096;
        CK(cudaMalloc(&l.d_raw, cap));
        l.raw_cap = cap;
    }
    if (blocks > l.blk_cap) {
        if (l.h_desc) cudaFreeHost(l.h_desc);
        if (l.d_desc) cudaFree(l.d_desc);
        if (l.h_status) cudaFreeHost(l.h_status);
        if (l.d_status) cudaFree(l.d_status);
        l.h_desc = nullptr; l.d_desc = nullptr; l.h_status = nullptr; l.d_status = nullptr;
        l.blk_cap = 0;
        const int cap = blocks + blocks / 4 + 16;
        CK(cudaMallocHost(&l.h_desc, cap * sizeof(Lz4BlockDesc)));
        CK(cudaMalloc(&l.d_desc, cap * sizeof(Lz4BlockDesc)));
        CK(cudaMallocHost(&l.h_status, cap * sizeof(int)));
        CK(cudaMalloc(&l.d_status, cap * sizeof(int)));
        l.blk_cap = cap;
    }
    return 0;
}

// wait for a lane's batch and validate what the decoder reported
int lz4_lane_retire(Lz4Lane& l)
{
    if (!l.busy) return 0;
    CK(cudaStreamSynchronize(l.st));
    l.busy = false;
    for (int b = 0; b < l.n; ++b)
        if (l.h_status[b] < 0 || (uint32_t)l.h_status[b] != l.h_desc[b].raw_size)
            return FLAGSTAT_CUDA_EFORMAT;
    return 0;
}

int lz4_lane_ship(int mode, int codec, Lz4Lane& l, size_t comp_bytes, size_t raw_bytes, bool all_even,
                  uint64_t* d_flags)
{
    if (l.n == 0) return 0;
    CK(cudaMemcpyAsync(l.d_comp, l.h_comp, comp_bytes, cudaMemcpyHostToDevice, l.st));
    CK(cudaMemcpyAsync(l.d_desc, l.h_desc, l.n * sizeof(Lz4BlockDesc), cudaMemcpyHostToDevice, l.st));
    TimerPair dbg;  // FLAGSTAT_CUDA_DEBUG: decode kernel time on stderr (events freed on every path)
    if (file_debug()) {
        CK(cudaEventCreate(&dbg.e0));
        CK(cudaEventCreate(&dbg.e1));
        CK(cudaEventRecord(dbg.e0, l.st));
    }
    int rc = codec == kCodecZstd
                 ? (zstd_variant() == 0
                        ? zstd_launch(l.d_comp, l.d_raw, l.d_desc, l.d_status, (uint32_t)l.n, l.d_work, l.st)
                        : zstd2_launch(l.z2, l.h_comp, l.h_desc, l.d_comp, l.d_raw, l.d_desc, l.d_status, (uint32_t)l.n, l.st))
                 : lz4_launch(l.d_comp, l.d_raw, l.d_desc, l.d_status, (uint32_t)l.n, l.st, l.scratch, l.max_comp,
                              l.max_raw);
    if (rc) return rc;
    if (file_debug()) {
        CK(cudaEventRecord(dbg.e1, l.st));
        CK(cudaEventSynchronize(dbg.e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, dbg.e0, dbg.e1);
        std::fprintf(stderr, "[flagstat_cuda] block decode: %d blocks, %zu -> %zu bytes, %.3f ms (%.1f GB/s out)\n",
                     l.n, comp_bytes, raw_bytes, ms, raw_bytes / (ms * 1e6));
    }
    if (all_even) {  // the decoded blocks are one contiguous run of whole records
        rc = launch(mode, reinterpret_cast<const uint16_t*>(l.d_raw), raw_bytes / 2, d_flags, l.st);
        if (rc) return rc;
    } else {         // a block with an odd byte count: the reference drops that byte (N = size >> 1)
        for (int b = 0; b < l.n; ++b) {
            rc = launch(mode, reinterpret_cast<const uint16_t*>(l.d_raw + l.h_desc[b].raw_off),
                        l.h_desc[b].raw_size / 2, d_flags, l.st);
            if (rc) return rc;
        }
    }
    CK(cudaMemcpyAsync(l.h_status, l.d_status, l.n * sizeof(int), cudaMemcpyDeviceToHost, l.st));
    l.busy = true;
    return 0;
}

constexpr int kLz4Lanes = 3;  // batch k is decoded while k + 1 crosses PCIe and k + 2 is gathered on the host
struct Lz4Ctx {
    Lz4Lane lanes[kLz4Lanes];
    uint64_t* d_flags = nullptr;
    int dev = -1;
};
std::mutex g_file_mu;
std::vector<Lz4Ctx*> g_lz4_pool[kMaxDevices];

int lz4_ctx_acquire(Lz4Ctx** out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_file_mu);
        if (!g_lz4_pool[dev].empty()) {
            *out = g_lz4_pool[dev].back();
            g_lz4_pool[dev].pop_back();
            return 0;
        }
    }
    Lz4Ctx* c = new (std::nothrow) Lz4Ctx();
    if (!c) return FLAGSTAT_CUDA_ENOMEM;
    c->dev = dev;
    const cudaError_t e = cudaMalloc(&c->d_flags, 32 * sizeof(uint64_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return (int)e;
    }
    *out = c;
    return 0;
}

void lz4_ctx_release(Lz4Ctx* c)
{
    for (auto& l : c->lanes) {
        if (l.st) cudaStreamSynchronize(l.st);
        l.busy = false;
        l.n = 0;
    }
    std::lock_guard<std::mutex> lk(g_file_mu);
    g_lz4_pool[c->dev].push_back(c);
}

// Walk the container's headers: [int32 raw][int32 comp][comp bytes] ... (flagstats.cpp:312-314)
int lz4_index(ByteSource& src, int fd, uint64_t total, std::vector<BlockInfo>& idx)
{
    uint64_t pos = 0;
    while (pos < total) {
        int32_t hdr[2];
        if (total - pos < sizeof(hdr)) return FLAGSTAT_CUDA_EFORMAT;
        if (src.mem) std::memcpy(hdr, src.mem + pos, sizeof(hdr));
        else if (::pread(fd, hdr, sizeof(hdr), (off_t)pos) != (ssize_t)sizeof(hdr)) return FLAGSTAT_CUDA_EIO;
        pos += sizeof(hdr);
        if (hdr[0] < 0 || hdr[1] <= 0 || (uint32_t)hdr[0] > kMaxRawBlock || (uint64_t)hdr[1] > total - pos)
            return FLAGSTAT_CUDA_EFORMAT;
        idx.push_back(BlockInfo{pos, (uint32_t)hdr[1], (uint32_t)hdr[0]});
        pos += (uint64_t)hdr[1];
    }
    return 0;
}

int consume_lz4(int mode, int codec, ByteSource& src, uint64_t* totals, uint64_t* n_records)
{
    int fd = -1;
    uint64_t total = src.size;
    if (src.fp) {
        struct stat sb;
        fd = ::fileno(src.fp);
        if (fd < 0 || ::fstat(fd, &sb) != 0) return FLAGSTAT_CUDA_EIO;
        total = (uint64_t)sb.st_size;
    }
    std::vector<BlockInfo> idx;
    int rc = lz4_index(src, fd, total, idx);
    if (rc) return rc;

    Lz4Ctx* ctx = nullptr;
    rc = lz4_ctx_acquire(&ctx);
    if (rc) return rc;
    struct Cleanup {
        Lz4Ctx* c;
        ~Cleanup() { lz4_ctx_release(c); }
    } cleanup{ctx};
    Lz4Lane* lanes = ctx->lanes;
    uint64_t* d_flags = ctx->d_flags;
    CK(cudaMemset(d_flags, 0, 32 * sizeof(uint64_t)));

    const int T = gather_threads();
    const bool nt = staging_nt();
    int batch_blocks = kBatchBlocks;
    if ((codec == kCodecLz4 && lz4_variant() == 2) || (codec == kCodecZstd && zstd_variant() != 0)) {
        // The CTA decoder works on 2 x SMs blocks at a time (one wave); batches of that size let the
        // gather and the H2D copy of batch k + 1 run while batch k is being decoded -- a file of
        // 1601 blocks used to be ONE batch: gather, copy, decode and count back to back.  A file of a
        // few hundred blocks is cut finer (one CTA per SM per batch: two batches decode side by side
        // and the first one starts after a 1.5 ms gather instead of 3 ms): 401 blocks 6.0 -> 5.3 ms at
        // ratio 5.2, 10.3 -> 8.1 ms at ratio 2.2; at 1601 blocks the wave-sized batch is the faster one.
        // Zstd frames: the entropy stage of a batch takes ~25 ms whatever the batch holds, so batches
        // are large (4 x SMs) and the stages of up to three of them overlap.
        // (profiles/r8d_container_batch_sweep.txt)
        int dev = 0;
        DeviceInfo* di = nullptr;
        CK(cudaGetDevice(&dev));
        const int irc = device_info(dev, &di);
        if (irc) return irc;
        if (codec == kCodecZstd) batch_blocks = 4 * di->sms;
        else batch_blocks = idx.size() <= (size_t)4 * (size_t)di->sms ? di->sms : 2 * di->sms;
        if (batch_blocks > kBatchBlocks) batch_blocks = kBatchBlocks;
    }
    if (const char* e = std::getenv("FLAGSTAT_CUDA_LZ4_BATCH")) {  // tests: force several batches
        const int v = std::atoi(e);
        if (v >= 1 && v <= kBatchBlocks) batch_blocks = v;
    }
    uint64_t records = 0;
    size_t first = 0;
    int cur = 0;
    while (first < idx.size()) {
        // batch [first, last): layout of the payloads (16-byte aligned) and of the decoded blocks
        size_t last = first, comp_bytes = 0, raw_bytes = 0;
        bool all_even = true;
        while (last < idx.size() && (int)(last - first) < batch_blocks) {
            const size_t r = idx[last].raw + (idx[last].raw & 1u);  // keep every block 2-byte aligned
            if (last > first && raw_bytes + r > kBatchRawCap) break;
            comp_bytes = ((comp_bytes + 15u) & ~(size_t)15u) + idx[last].comp;
            raw_bytes += r;
            all_even = all_even && (idx[last].raw & 1u) == 0u;
            ++last;
        }
        Lz4Lane& l = lanes[cur];
        rc = lz4_lane_retire(l);  // its previous batch must be off the staging buffers
        if (rc) return rc;
        rc = lz4_lane_reserve(l, comp_bytes, raw_bytes, (int)(last - first), codec);
        if (rc) return rc;
        l.n = (int)(last - first);
        l.max_comp = l.max_raw = 0;
        size_t c = 0, r = 0;
        for (size_t b = first; b < last; ++b) {
            c = (c + 15u) & ~(size_t)15u;
            if (idx[b].comp > l.max_comp) l.max_comp = idx[b].comp;
            if (idx[b].raw > l.max_raw) l.max_raw = idx[b].raw;
            l.h_desc[b - first] = Lz4BlockDesc{c, r, idx[b].comp, idx[b].raw};
            c += idx[b].comp;
            r += idx[b].raw + (idx[b].raw & 1u);
            records += idx[b].raw >> 1;
        }
        // gather the payloads into the pinned staging buffer with T threads
        std::atomic<int> err{0};
        auto gather = [&](int t) {
            for (size_t b = first + (size_t)t; b < last; b += (size_t)T) {
                unsigned char* dst = l.h_comp + l.h_desc[b - first].comp_off;
                if (src.mem) {
                    if (nt) copy_streaming(dst, src.mem + idx[b].src_off, idx[b].comp);
                    else std::memcpy(dst, src.mem + idx[b].src_off, idx[b].comp);
                } else {
                    size_t got = 0;
                    while (got < idx[b].comp) {
                        const ssize_t k = ::pread(fd, dst + got, idx[b].comp - got, (off_t)(idx[b].src_off + got));
                        if (k <= 0) {
                            err.store(FLAGSTAT_CUDA_EIO);
                            return;
                        }
                        got += (size_t)k;
                    }
                }
            }
        };
        {
            std::vector<std::thread> th;
            const int use = (l.n < 4 * T) ? 1 : T;  // not worth threads for a handful of blocks
            if (use == 1) {
                for (int t = 0; t < T; ++t) gather(t);
            } else {
                for (int t = 1; t < T; ++t) th.emplace_back(gather, t);
                gather(0);
                for (auto& x : th) x.join();
            }
        }
        if (err.load()) return err.load();
        rc = lz4_lane_ship(mode, codec, l, c, r, all_even, d_flags);
        if (rc) return rc;
        cur = (cur + 1) % kLz4Lanes;
        first = last;
    }
    for (int i = 0; i < kLz4Lanes; ++i) {
        rc = lz4_lane_retire(lanes[i]);
        if (rc) return rc;
    }
    CK(cudaMemcpy(totals, d_flags, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *n_records = records;
    return 0;
}

// Raw ".bin" files: T reader threads, each with two private ring slots (8 blocks = 8,192,000
// bytes each: the coalesced-DMA size of the block stream).  A thread pread()s its next
// group straight into a pinned slot, enqueues one DMA + one kernel on the slot's stream and
// moves to its other slot, so the page-cache copy of group k+1 overlaps DMA + count of group
// k, and T such pipelines run side by side (one reader thread tops out at ~5 GB/s).
constexpr size_t kRawSlotBytes = 8u * (size_t)kRefBlockBytes;
constexpr int kRawMaxThreads = 24;

struct RawCtx {
    int dev = -1;
    int threads = 0;
    unsigned char* h = nullptr;  // pinned, threads * 2 slots
    unsigned char* d = nullptr;
    cudaStream_t st[2 * kRawMaxThreads] = {};
    uint64_t* d_flags = nullptr;
};
std::vector<RawCtx*> g_raw_pool[kMaxDevices];

int raw_ctx_acquire(RawCtx** out)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    const int want = io_threads();
    {
        std::lock_guard<std::mutex> lk(g_file_mu);
        auto& pool = g_raw_pool[dev];
        for (size_t i = 0; i < pool.size(); ++i)
            if (pool[i]->threads == want) {
                *out = pool[i];
                pool.erase(pool.begin() + (long)i);
                return 0;
            }
    }
    RawCtx* c = new (std::nothrow) RawCtx();
    if (!c) return FLAGSTAT_CUDA_ENOMEM;
    c->dev = dev;
    c->threads = want;
    auto build = [&]() -> int {
        CK(cudaHostAlloc(&c->h, (size_t)want * 2 * kRawSlotBytes, cudaHostAllocPortable));
        CK(cudaMalloc(&c->d, (size_t)want * 2 * kRawSlotBytes));
        for (int i = 0; i < 2 * want; ++i) CK(cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking));
        CK(cudaMalloc(&c->d_flags, 32 * sizeof(uint64_t)));
        return 0;
    };
    const int rc = build();
    if (rc) {  // e.g. not enough pinned memory: give back what was taken
        cudaGetLastError();
        for (int i = 0; i < 2 * want; ++i)
            if (c->st[i]) cudaStreamDestroy(c->st[i]);
        if (c->h) cudaFreeHost(c->h);
        if (c->d) cudaFree(c->d);
        if (c->d_flags) cudaFree(c->d_flags);
        delete c;
        return rc;
    }
    *out = c;
    return 0;
}

void raw_ctx_release(RawCtx* c)
{
    std::lock_guard<std::mutex> lk(g_file_mu);
    g_raw_pool[c->dev].push_back(c);
}

// The staged pipeline shared by raw files and pageable host arrays: `fill(dst, off, len)` puts
// bytes [off, off + len) of the column into a pinned slot (pread for files, memcpy for memory).
// slot_bytes <= kRawSlotBytes, even.  totals: 32 (flagstat) / 16 (pospopcnt) u64, overwritten.
template <class Fill>
int consume_staged(int mode, uint64_t size, size_t slot_bytes, Fill fill, uint64_t* totals)
{
    RawCtx* c = nullptr;
    {
        const int rc = raw_ctx_acquire(&c);
        if (rc) return rc;
    }
    struct Cleanup {
        RawCtx* c;
        ~Cleanup() { raw_ctx_release(c); }
    } cleanup{c};
    CK(cudaMemset(c->d_flags, 0, 32 * sizeof(uint64_t)));
    const uint64_t n_groups = (size + slot_bytes - 1) / slot_bytes;
    std::atomic<int> err{0};
    auto worker = [&](int t) {
        if (cudaSetDevice(c->dev) != cudaSuccess) {
            err.store(FLAGSTAT_CUDA_ENODEV);
            return;
        }
        int k = 0;
        for (uint64_t g = (uint64_t)t; g < n_groups && err.load() == 0; g += (uint64_t)c->threads, ++k) {
            const int slot = 2 * t + (k & 1);
            unsigned char* h = c->h + (size_t)slot * kRawSlotBytes;
            unsigned char* d = c->d + (size_t)slot * kRawSlotBytes;
            const uint64_t off = g * slot_bytes;
            const size_t len = (size_t)((size - off < slot_bytes) ? (size - off) : slot_bytes);
            if (cudaStreamSynchronize(c->st[slot]) != cudaSuccess) {  // slot's previous DMA is done
                err.store(FLAGSTAT_CUDA_EIO);
                return;
            }
            if (!fill(h, off, len)) {
                err.store(FLAGSTAT_CUDA_EIO);
                return;
            }
            const uint64_t recs = len >> 1;  // only the last group can carry an odd byte; dropped like :455
            if (recs == 0) continue;
            if (cudaMemcpyAsync(d, h, recs * 2, cudaMemcpyHostToDevice, c->st[slot]) != cudaSuccess) {
                err.store(FLAGSTAT_CUDA_EIO);
                return;
            }
            const int rc = launch(mode, reinterpret_cast<const uint16_t*>(d), recs, c->d_flags, c->st[slot]);
            if (rc) {
                err.store(rc);
                return;
            }
        }
    };
    const int use = (int)((n_groups < (uint64_t)c->threads) ? n_groups : (uint64_t)c->threads);
    std::vector<std::thread> th;
    for (int t = 1; t < use; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& x : th) x.join();
    for (int i = 0; i < 2 * c->threads; ++i) CK(cudaStreamSynchronize(c->st[i]));
    if (err.load()) return err.load();
    CK(cudaMemcpy(totals, c->d_flags, (mode == kPospopcnt ? 16 : 32) * sizeof(uint64_t),
                  cudaMemcpyDeviceToHost));
    return 0;
}

// Slot size of the raw-file reader: 8 MB DMAs reach the link rate when ONE producer feeds them
// (block stream, section 8 of DESIGN.md), but T readers that all start with an 8 MB pread leave
// the link idle for the first ~0.8 ms and drain as long at the end -- a tenth of a 410 MB file.
// Smaller slots start the first DMA earlier; the per-DMA fixed cost overlaps across the T
// streams.  FLAGSTAT_CUDA_RAW_SLOT_KB overrides (A/B: tools/file_bench.py).
size_t raw_slot_bytes(uint64_t file_size, int threads)
{
    if (const char* e = std::getenv("FLAGSTAT_CUDA_RAW_SLOT_KB")) {
        size_t v = (size_t)std::strtoull(e, nullptr, 10) << 10;
        v &= ~(size_t)65535u;
        if (v >= (64u << 10) && v <= kRawSlotBytes) return v;
    }
    // measured (profiles/r4j_raw_reader_sweep.jsonl, 12 readers): 4 MiB slots 47.9 GB/s on the 1.65 GB
    // column (0.94 of the pinned-array rate of that box), 2 MiB 44.4, 8 MB 44.2, 1 MiB 37.7;
    // on a 410 MB file 39.4 / 38.4 / 35.5 / 34.3.  Short files get at least ~4 slots per reader.
    size_t v = 4u << 20;
    const size_t share = (size_t)(file_size / ((uint64_t)threads * 4u));
    if (share < v) v = (share + 65535u) & ~(size_t)65535u;
    if (v < (1u << 20)) v = 1u << 20;
    return v;
}

int consume_raw_fd(int mode, int fd, uint64_t size, uint64_t* totals, uint64_t* n_records)
{
    const int rc = consume_staged(
        mode, size, raw_slot_bytes(size, io_threads()),
        [fd](unsigned char* h, uint64_t off, size_t len) {
            size_t got = 0;
            while (got < len) {
                const ssize_t r = ::pread(fd, h + got, len - got, (off_t)(off + got));
                if (r <= 0) return false;
                got += (size_t)r;
            }
            return true;
        },
        totals);
    if (rc) return rc;
    *n_records = size >> 1;
    return 0;
}

// Pageable (unregistered) host arrays -- what numpy / malloc hand the reference's entry points.
// cudaMemcpyAsync from such memory is staged by the driver through its own bounce buffer on
// the calling thread; here T threads memcpy their slices into pinned slots and every slice is
// its own DMA + launch, so the page-in copy of one slice overlaps the DMA of the others.
// (Pinning the caller's pages in place instead -- cudaHostRegister per 32 MiB granule, DMA straight
// from user memory, unregister behind it -- was built and measured: 8.6-9.6 GB/s whatever the
// thread count, the per-granule pin / unpin calls serialise in the driver; one whole-array
// cudaHostRegister costs 36-55 ms + 27-40 ms to undo per 1.65 GB.  profiles/r4e_pageable.jsonl.)
int run_pageable(int mode, const uint16_t* array, uint64_t len, uint64_t* totals)
{
    const uint64_t size = len * sizeof(uint16_t);
    const int T = io_threads();
    // at least two slices per thread, 64 KiB granules, at most one ring slot
    uint64_t slot = (size / (uint64_t)(2 * T) + 65535u) & ~(uint64_t)65535u;
    if (slot < (1u << 20)) slot = 1u << 20;
    if (slot > kRawSlotBytes) slot = kRawSlotBytes & ~(size_t)65535u;
    const unsigned char* base = reinterpret_cast<const unsigned char*>(array);
    const bool nt = staging_nt();
    return consume_staged(
        mode, size, (size_t)slot,
        [base, nt](unsigned char* h, uint64_t off, size_t n) {
            if (nt) copy_streaming(h, base + off, n);
            else std::memcpy(h, base + off, n);
            return true;
        },
        totals);
}

int consume_raw(int mode, ByteSource& src, uint64_t* totals, uint64_t* n_records)
{
    if (src.fp) {
        struct stat sb;
        const int fd = ::fileno(src.fp);
        if (fd < 0 || ::fstat(fd, &sb) != 0) return FLAGSTAT_CUDA_EIO;
        return consume_raw_fd(mode, fd, (uint64_t)sb.st_size, totals, n_records);
    }
    // already in host memory: the host-pointer path of FLAGSTAT_cuda_u64 (chunked, overlapped staging)
    *n_records = src.size >> 1;
    return run_sync(mode, reinterpret_cast<const uint16_t*>(src.mem), src.size >> 1, totals);
}

int consume_impl(ByteSource& src, int format, uint64_t* flags, uint64_t* n_records);
int consume(ByteSource& src, int format, uint64_t* flags, uint64_t* n_records)
{
    return guarded([&] { return consume_impl(src, format, flags, n_records); });
}

int consume_impl(ByteSource& src, int format, uint64_t* flags, uint64_t* n_records)
{
    if (!flags) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    uint64_t t[32] = {0}, n = 0;
    int rc;
    // _SAMTOOLS: the reference's "samtools" callers of the same files (flagstat_loop over every
    // block, benchmark/flagstats.cpp:496-519 raw, :547-590 LZ4): counters + exact n_pair_all
    const int mode = (format & FLAGSTAT_CUDA_FILE_SAMTOOLS) ? kSamtools : kFlagstat;
    format &= ~FLAGSTAT_CUDA_FILE_SAMTOOLS;
    if (format == FLAGSTAT_CUDA_FILE_RAW) rc = consume_raw(mode, src, t, &n);
    else if (format == FLAGSTAT_CUDA_FILE_LZ4) rc = consume_lz4(mode, kCodecLz4, src, t, &n);
    else if (format == FLAGSTAT_CUDA_FILE_ZSTD) rc = consume_lz4(mode, kCodecZstd, src, t, &n);
    else return FLAGSTAT_CUDA_EINVAL;
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) flags[i] += t[i];
    if (n_records) *n_records = n;
    return 0;
}

}  // namespace

extern "C" {

int FLAGSTAT_cuda_set_lz4_variant(int v)
{
    const int prev = lz4_variant();
    g_lz4_variant.store((v < 0 || v > 2) ? 2 : v);
    return prev;
}

int FLAGSTAT_cuda_file_u64(const char* path, int format, uint64_t* flags, uint64_t* n_records)
{
    if (!path) return FLAGSTAT_CUDA_EINVAL;
    ByteSource src;
    src.fp = std::fopen(path, "rb");
    if (!src.fp) return FLAGSTAT_CUDA_EIO;
    const int rc = consume(src, format, flags, n_records);  // every reader pread()s the descriptor
    std::fclose(src.fp);
    return rc;
}

int FLAGSTAT_cuda_container_u64(const void* bytes, uint64_t n_bytes, int format, uint64_t* flags,
                                uint64_t* n_records)
{
    if (!bytes && n_bytes) return FLAGSTAT_CUDA_EINVAL;
    ByteSource src;
    src.mem = static_cast<const unsigned char*>(bytes);
    src.size = n_bytes;
    return consume(src, format, flags, n_records);
}

// Decode-only entries (tests, tools): n_blocks payloads described by host arrays; the decoded
// bytes are returned in `raw` (host, raw_total bytes).  status[b] = decoded size or < 0.
static int decode_blocks(int codec, const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                         const uint32_t* comp_size, const uint64_t* raw_off, const uint32_t* raw_size,
                         uint32_t n_blocks, void* raw, uint64_t raw_total, int* status)
{
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (n_blocks == 0) return 0;
    if (!comp || !raw || !status || !comp_off || !comp_size || !raw_off || !raw_size) return FLAGSTAT_CUDA_EINVAL;
    return guarded([&]() -> int {
    std::vector<Lz4BlockDesc> desc(n_blocks);
    for (uint32_t b = 0; b < n_blocks; ++b) {
        // written so that an offset near 2^64 cannot wrap the sum past the bound
        if (comp_off[b] > comp_bytes || comp_size[b] > comp_bytes - comp_off[b] || raw_off[b] > raw_total ||
            raw_size[b] > raw_total - raw_off[b])
            return FLAGSTAT_CUDA_EINVAL;
        desc[b] = Lz4BlockDesc{comp_off[b], raw_off[b], comp_size[b], raw_size[b]};
    }
    unsigned char *d_comp = nullptr, *d_raw = nullptr;
    Lz4BlockDesc* d_desc = nullptr;
    int* d_status = nullptr;
    zstd::Work* d_work = nullptr;
    Zstd2 z2;
    Lz4Scratch scratch;
    uint32_t max_comp = 0, max_raw = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (comp_size[b] > max_comp) max_comp = comp_size[b];
        if (raw_size[b] > max_raw) max_raw = raw_size[b];
    }
    int rc = 0;
    do {
        if ((rc = (int)cudaMalloc(&d_comp, comp_bytes ? comp_bytes : 1))) break;
        if ((rc = (int)cudaMalloc(&d_raw, raw_total ? raw_total : 1))) break;
        if ((rc = (int)cudaMalloc(&d_desc, n_blocks * sizeof(Lz4BlockDesc)))) break;
        if ((rc = (int)cudaMalloc(&d_status, n_blocks * sizeof(int)))) break;
        if ((rc = (int)cudaMemcpy(d_comp, comp, comp_bytes, cudaMemcpyHostToDevice))) break;
        if ((rc = (int)cudaMemset(d_raw, 0, raw_total ? raw_total : 1))) break;
        if ((rc = (int)cudaMemcpy(d_desc, desc.data(), n_blocks * sizeof(Lz4BlockDesc), cudaMemcpyHostToDevice))) break;
        if (codec == kCodecZstd && zstd_variant() != 0) {
            if ((rc = zstd2_launch(z2, static_cast<const unsigned char*>(comp), desc.data(), d_comp, d_raw, d_desc,
                                   d_status, n_blocks, nullptr)))
                break;
        } else if (codec == kCodecZstd) {
            if ((rc = (int)cudaMalloc(&d_work, (size_t)n_blocks * sizeof(zstd::Work)))) break;
            if ((rc = zstd_launch(d_comp, d_raw, d_desc, d_status, n_blocks, d_work, nullptr))) break;
        } else if ((rc = lz4_launch(d_comp, d_raw, d_desc, d_status, n_blocks, nullptr, scratch, max_comp, max_raw))) {
            break;
        }
        if ((rc = (int)cudaMemcpy(raw, d_raw, raw_total, cudaMemcpyDeviceToHost))) break;
        if ((rc = (int)cudaMemcpy(status, d_status, n_blocks * sizeof(int), cudaMemcpyDeviceToHost))) break;
    } while (0);
    cudaFree(d_comp);
    cudaFree(d_raw);
    cudaFree(d_desc);
    cudaFree(d_status);
    cudaFree(d_work);
    if (rc == 0 || z2.h_aux) cudaDeviceSynchronize();  // (the aux upload is asynchronous: nothing may be in flight when it is freed)
    zstd2_free(z2);
    cudaFree(scratch.d);
    return rc;
    });
}

int FLAGSTAT_cuda_lz4_decode(const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                             const uint32_t* comp_size, const uint64_t* raw_off, const uint32_t* raw_size,
                             uint32_t n_blocks, void* raw, uint64_t raw_total, int* status)
{
    return decode_blocks(kCodecLz4, comp, comp_bytes, comp_off, comp_size, raw_off, raw_size, n_blocks, raw,
                         raw_total, status);
}

int FLAGSTAT_cuda_zstd_decode(const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                              const uint32_t* comp_size, const uint64_t* raw_off, const uint32_t* raw_size,
                              uint32_t n_blocks, void* raw, uint64_t raw_total, int* status)
{
    return decode_blocks(kCodecZstd, comp, comp_bytes, comp_off, comp_size, raw_off, raw_size, n_blocks, raw,
                         raw_total, status);
}

}  // extern "C"
