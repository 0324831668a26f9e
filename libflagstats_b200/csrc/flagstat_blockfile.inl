// flagstat_blockfile.inl -- consumer of the reference's FLAG files (included by
// flagstat_capi.cu inside extern "C").
//
//   FLAGSTAT_CUDA_FILE_RAW  plain uint16 stream (".bin"), which the reference reads
//                           in 1,024,000-byte blocks (benchmark/flagstats.cpp:415-468)
//   FLAGSTAT_CUDA_FILE_LZ4  [int32 raw_size][int32 comp_size][LZ4 block] records
//                           (written :110-186, read :288-358)
//
// Raw files go through the pinned block ring (file -> pinned slot -> DMA ->
// kernel).  LZ4 containers are shipped COMPRESSED: payloads are gathered into a
// pinned staging buffer, one DMA per batch of blocks, lz4_decode_kernel (one
// warp per block) expands them in HBM and the flagstat kernel reads the decoded
// records from there.  Two batches are in flight (host gather of batch k+1
// overlaps DMA + decode + count of batch k).

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
constexpr int kBatchBlocks = 128;
constexpr size_t kBatchRawCap = (size_t)kBatchBlocks * kRefBlockBytes;
constexpr size_t kBatchCompCap = kBatchRawCap + (kBatchRawCap / 255) + 16u * kBatchBlocks + 4096u;

struct Lz4Lane {
    unsigned char* h_comp = nullptr;   // pinned
    unsigned char* d_comp = nullptr;
    unsigned char* d_raw = nullptr;
    Lz4BlockDesc* h_desc = nullptr;    // pinned
    Lz4BlockDesc* d_desc = nullptr;
    int* h_status = nullptr;           // pinned
    int* d_status = nullptr;
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
    if (l.st) cudaStreamDestroy(l.st);
    l = Lz4Lane();
}

int lz4_lane_alloc(Lz4Lane& l)
{
    CK(cudaMallocHost(&l.h_comp, kBatchCompCap));
    CK(cudaMalloc(&l.d_comp, kBatchCompCap));
    CK(cudaMalloc(&l.d_raw, kBatchRawCap + 256));  // + one pad byte per odd-sized block
    CK(cudaMallocHost(&l.h_desc, kBatchBlocks * sizeof(Lz4BlockDesc)));
    CK(cudaMalloc(&l.d_desc, kBatchBlocks * sizeof(Lz4BlockDesc)));
    CK(cudaMallocHost(&l.h_status, kBatchBlocks * sizeof(int)));
    CK(cudaMalloc(&l.d_status, kBatchBlocks * sizeof(int)));
    CK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
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

int lz4_lane_ship(Lz4Lane& l, size_t comp_bytes, size_t raw_bytes, bool all_even, uint64_t* d_flags)
{
    if (l.n == 0) return 0;
    CK(cudaMemcpyAsync(l.d_comp, l.h_comp, comp_bytes, cudaMemcpyHostToDevice, l.st));
    CK(cudaMemcpyAsync(l.d_desc, l.h_desc, l.n * sizeof(Lz4BlockDesc), cudaMemcpyHostToDevice, l.st));
    const unsigned grid = (unsigned)((l.n + kLz4WarpsPerCta - 1) / kLz4WarpsPerCta);
    lz4_decode_kernel<<<grid, kLz4WarpsPerCta * 32, 0, l.st>>>(l.d_comp, l.d_raw, l.d_desc, l.d_status,
                                                              (uint32_t)l.n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
    if (all_even) {  // the decoded blocks are one contiguous run of whole records
        const int rc = launch(kFlagstat, reinterpret_cast<const uint16_t*>(l.d_raw), raw_bytes / 2, d_flags, l.st);
        if (rc) return rc;
    } else {         // a block with an odd byte count: the reference drops that byte (N = size >> 1)
        for (int b = 0; b < l.n; ++b) {
            const int rc = launch(kFlagstat, reinterpret_cast<const uint16_t*>(l.d_raw + l.h_desc[b].raw_off),
                                  l.h_desc[b].raw_size / 2, d_flags, l.st);
            if (rc) return rc;
        }
    }
    CK(cudaMemcpyAsync(l.h_status, l.d_status, l.n * sizeof(int), cudaMemcpyDeviceToHost, l.st));
    l.busy = true;
    return 0;
}

int consume_lz4(ByteSource& src, uint64_t* totals, uint64_t* n_records)
{
    Lz4Lane lanes[2];
    uint64_t* d_flags = nullptr;
    struct Cleanup {
        Lz4Lane* l;
        uint64_t** f;
        ~Cleanup()
        {
            lz4_lane_free(l[0]);
            lz4_lane_free(l[1]);
            if (*f) cudaFree(*f);
        }
    } cleanup{lanes, &d_flags};
    for (auto& l : lanes) {
        const int rc = lz4_lane_alloc(l);
        if (rc) return rc;
    }
    CK(cudaMalloc(&d_flags, 32 * sizeof(uint64_t)));
    CK(cudaMemset(d_flags, 0, 32 * sizeof(uint64_t)));
    uint64_t records = 0;
    int cur = 0;
    bool done = false;
    while (!done) {
        Lz4Lane& l = lanes[cur];
        int rc = lz4_lane_retire(l);  // its previous batch must be off the staging buffers
        if (rc) return rc;
        l.n = 0;
        size_t comp_bytes = 0, raw_bytes = 0;
        bool all_even = true;
        while (l.n < kBatchBlocks) {
            if (src.at_end()) {
                done = true;
                break;
            }
            int32_t hdr[2];
            if (src.read(hdr, sizeof(hdr)) != sizeof(hdr)) return FLAGSTAT_CUDA_EFORMAT;
            if (hdr[0] < 0 || hdr[1] <= 0 || (uint32_t)hdr[0] > kMaxRawBlock ||
                (uint32_t)hdr[1] > kMaxRawBlock + kMaxRawBlock / 255 + 16)
                return FLAGSTAT_CUDA_EFORMAT;
            const size_t comp_at = (comp_bytes + 15u) & ~(size_t)15u;
            if (comp_at + (size_t)hdr[1] > kBatchCompCap || raw_bytes + (size_t)hdr[0] > kBatchRawCap) {
                if (l.n == 0) return FLAGSTAT_CUDA_EFORMAT;  // one block larger than a whole batch
                // does not fit: rewind the header and close the batch
                if (src.fp) std::fseek(src.fp, -(long)sizeof(hdr), SEEK_CUR);
                src.pos -= sizeof(hdr);
                break;
            }
            if (src.read(l.h_comp + comp_at, (size_t)hdr[1]) != (size_t)hdr[1]) return FLAGSTAT_CUDA_EFORMAT;
            Lz4BlockDesc& d = l.h_desc[l.n];
            d.comp_off = comp_at;
            d.comp_size = (uint32_t)hdr[1];
            d.raw_off = raw_bytes;
            d.raw_size = (uint32_t)hdr[0];
            comp_bytes = comp_at + (size_t)hdr[1];
            raw_bytes += (size_t)hdr[0];
            if (hdr[0] & 1) {
                all_even = false;
                raw_bytes += 1;  // keep every block's first record 2-byte aligned
            }
            records += (uint64_t)hdr[0] >> 1;
            ++l.n;
        }
        rc = lz4_lane_ship(l, comp_bytes, raw_bytes, all_even, d_flags);
        if (rc) return rc;
        cur ^= 1;
    }
    for (auto& l : lanes) {
        const int rc = lz4_lane_retire(l);
        if (rc) return rc;
    }
    CK(cudaMemcpy(totals, d_flags, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *n_records = records;
    return 0;
}

int consume_raw(ByteSource& src, uint64_t* totals, uint64_t* n_records)
{
    int dev = 0;
    CK(cudaGetDevice(&dev));
    FLAGSTAT_cuda_stream* s = nullptr;
    int rc = FLAGSTAT_cuda_stream_open_ex(&s, dev, kRefBlockBytes / 2, 4, FLAGSTAT_CUDA_STREAM_DMA, 8);
    if (rc) return rc;
    uint64_t records = 0;
    for (;;) {
        uint16_t* slot = FLAGSTAT_cuda_stream_acquire(s);
        if (!slot) {
            rc = FLAGSTAT_CUDA_ESTATE;
            break;
        }
        const size_t got = src.read(slot, kRefBlockBytes);
        rc = FLAGSTAT_cuda_stream_submit(s, (uint32_t)(got >> 1));  // an odd trailing byte is dropped, like :455
        if (rc) break;
        records += got >> 1;
        if (got < kRefBlockBytes) break;
    }
    if (!rc) {
        for (int i = 0; i < 32; ++i) totals[i] = 0;
        rc = FLAGSTAT_cuda_stream_finish(s, totals);
    }
    FLAGSTAT_cuda_stream_close(s);
    *n_records = records;
    return rc;
}

int consume(ByteSource& src, int format, uint64_t* flags, uint64_t* n_records)
{
    if (!flags) return FLAGSTAT_CUDA_EINVAL;
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    uint64_t t[32] = {0}, n = 0;
    int rc;
    if (format == FLAGSTAT_CUDA_FILE_RAW) rc = consume_raw(src, t, &n);
    else if (format == FLAGSTAT_CUDA_FILE_LZ4) rc = consume_lz4(src, t, &n);
    else return FLAGSTAT_CUDA_EINVAL;
    if (rc) return rc;
    for (int i = 0; i < 32; ++i) flags[i] += t[i];
    if (n_records) *n_records = n;
    return 0;
}

}  // namespace

extern "C" {

int FLAGSTAT_cuda_file_u64(const char* path, int format, uint64_t* flags, uint64_t* n_records)
{
    if (!path) return FLAGSTAT_CUDA_EINVAL;
    ByteSource src;
    src.fp = std::fopen(path, "rb");
    if (!src.fp) return FLAGSTAT_CUDA_EIO;
    static thread_local std::vector<char> iobuf;
    iobuf.resize(4u << 20);
    std::setvbuf(src.fp, iobuf.data(), _IOFBF, iobuf.size());
    const int rc = consume(src, format, flags, n_records);
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

// Decode-only entry (tests, tools): n_blocks LZ4 blocks described by host arrays; the decoded
// bytes are returned in `raw` (host, raw_total bytes).  status[b] = decoded size or < 0.
int FLAGSTAT_cuda_lz4_decode(const void* comp, uint64_t comp_bytes, const uint64_t* comp_off,
                             const uint32_t* comp_size, const uint64_t* raw_off, const uint32_t* raw_size,
                             uint32_t n_blocks, void* raw, uint64_t raw_total, int* status)
{
    if (probe_devices() <= 0) return FLAGSTAT_CUDA_ENODEV;
    if (n_blocks == 0) return 0;
    if (!comp || !raw || !status) return FLAGSTAT_CUDA_EINVAL;
    std::vector<Lz4BlockDesc> desc(n_blocks);
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (comp_off[b] + comp_size[b] > comp_bytes || raw_off[b] + raw_size[b] > raw_total)
            return FLAGSTAT_CUDA_EINVAL;
        desc[b] = Lz4BlockDesc{comp_off[b], raw_off[b], comp_size[b], raw_size[b]};
    }
    unsigned char *d_comp = nullptr, *d_raw = nullptr;
    Lz4BlockDesc* d_desc = nullptr;
    int* d_status = nullptr;
    int rc = 0;
    do {
        if ((rc = (int)cudaMalloc(&d_comp, comp_bytes ? comp_bytes : 1))) break;
        if ((rc = (int)cudaMalloc(&d_raw, raw_total ? raw_total : 1))) break;
        if ((rc = (int)cudaMalloc(&d_desc, n_blocks * sizeof(Lz4BlockDesc)))) break;
        if ((rc = (int)cudaMalloc(&d_status, n_blocks * sizeof(int)))) break;
        if ((rc = (int)cudaMemcpy(d_comp, comp, comp_bytes, cudaMemcpyHostToDevice))) break;
        if ((rc = (int)cudaMemset(d_raw, 0, raw_total ? raw_total : 1))) break;
        if ((rc = (int)cudaMemcpy(d_desc, desc.data(), n_blocks * sizeof(Lz4BlockDesc), cudaMemcpyHostToDevice))) break;
        const unsigned grid = (n_blocks + kLz4WarpsPerCta - 1) / kLz4WarpsPerCta;
        lz4_decode_kernel<<<grid, kLz4WarpsPerCta * 32>>>(d_comp, d_raw, d_desc, d_status, n_blocks);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if ((rc = (int)cudaGetLastError())) break;
        if ((rc = (int)cudaMemcpy(raw, d_raw, raw_total, cudaMemcpyDeviceToHost))) break;
        if ((rc = (int)cudaMemcpy(status, d_status, n_blocks * sizeof(int), cudaMemcpyDeviceToHost))) break;
    } while (0);
    cudaFree(d_comp);
    cudaFree(d_raw);
    cudaFree(d_desc);
    cudaFree(d_status);
    return rc;
}

}  // extern "C"
