// h2d_locality.cu -- follow-up to h2d_matrix.cu.  The matrix (profiles/r4a_h2d_matrix.jsonl)
// showed: every GPU alone and every PAIR of GPUs pulls pinned host memory at 55.6 GB/s, GPUs
// {4,5,6,7} together 4 x 55, but GPUs {0,1,2,3} together only 4 x 28.9 = 115 GB/s, and all
// eight 4 x 23.4 + 4 x 35.5 = 235 GB/s -- the signature of host memory that sits behind ONE
// socket (115 GB/s = what crosses the inter-socket links, 235 GB/s = one socket's DRAM), in a
// guest that sees a single NUMA node.  Question answered here: does it depend on WHERE in guest
// memory the pinned buffer lies?  K chunks of C MiB are allocated (cudaHostAlloc), their guest
// physical address is read from /proc/self/pagemap, and for every chunk the GPUs of group A
// (first half) and then of group B (second half) copy disjoint slices of it concurrently.  A
// chunk that is local to a group's socket shows ~4 x 55 GB/s there and ~115 for the other group.
//
//   tools/bin/h2d_locality [chunks = 48] [MiB per chunk = 1024]      -> JSON lines
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            std::exit(2);                                                                      \
        }                                                                                      \
    } while (0)

static unsigned long long gpa_of(const void* p)
{
    static int fd = open("/proc/self/pagemap", O_RDONLY);
    if (fd < 0) return 0;
    unsigned long long e = 0;
    const off_t off = (off_t)((uintptr_t)p / 4096u * 8u);
    if (pread(fd, &e, 8, off) != 8) return 0;
    if (!(e >> 63)) return 0;                      // not present
    return (e & ((1ull << 55) - 1)) * 4096ull;     // PFN -> guest physical address
}

struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

struct Gpu { int dev; void* d; cudaStream_t st; cudaEvent_t e0, e1; };

// GPUs `sel` copy slice i of `chunk` (slice bytes each) concurrently; returns aggregate GB/s (wall)
static double group_rate(std::vector<Gpu>& gpus, const std::vector<int>& sel, const char* chunk, size_t slice,
                         int reps, std::vector<double>* per)
{
    const int k = (int)sel.size();
    Barrier bar(k + 1);
    std::vector<float> ms(k, 0.f);
    std::vector<std::thread> th;
    for (int i = 0; i < k; ++i)
        th.emplace_back([&, i] {
            Gpu& g = gpus[sel[i]];
            CK(cudaSetDevice(g.dev));
            const char* src = chunk + (size_t)i * slice;
            bar.wait();
            CK(cudaEventRecord(g.e0, g.st));
            for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(g.d, src, slice, cudaMemcpyHostToDevice, g.st));
            CK(cudaEventRecord(g.e1, g.st));
            CK(cudaStreamSynchronize(g.st));
            bar.wait();
            CK(cudaEventElapsedTime(&ms[i], g.e0, g.e1));
        });
    bar.wait();
    const auto t0 = std::chrono::steady_clock::now();
    bar.wait();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& t : th) t.join();
    if (per) {
        per->clear();
        for (int i = 0; i < k; ++i) per->push_back((double)slice * reps / (ms[i] * 1e-3) / 1e9);
    }
    return (double)slice * reps * k / wall / 1e9;
}

int main(int argc, char** argv)
{
    const int chunks = argc > 1 ? std::atoi(argv[1]) : 48;
    const size_t mib = argc > 2 ? (size_t)std::atol(argv[2]) : 1024;
    const size_t bytes = mib << 20;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 4) { std::printf("{\"note\": \"needs >= 4 GPUs\"}\n"); return 0; }
    const int half = ndev / 2;
    const size_t slice = bytes / (size_t)half;
    std::vector<Gpu> gpus(ndev);
    for (int g = 0; g < ndev; ++g) {
        gpus[g].dev = g;
        CK(cudaSetDevice(g));
        CK(cudaMalloc(&gpus[g].d, slice));
        CK(cudaStreamCreateWithFlags(&gpus[g].st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&gpus[g].e0));
        CK(cudaEventCreate(&gpus[g].e1));
    }
    std::vector<int> A, B;
    for (int g = 0; g < half; ++g) A.push_back(g);
    for (int g = half; g < ndev; ++g) B.push_back(g);
    CK(cudaSetDevice(0));
    std::vector<char*> buf(chunks, nullptr);
    for (int c = 0; c < chunks; ++c) {
        const auto t0 = std::chrono::steady_clock::now();
        CK(cudaHostAlloc((void**)&buf[c], bytes, cudaHostAllocPortable));
        std::memset(buf[c], c + 1, bytes);
        const double alloc_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const unsigned long long g0 = gpa_of(buf[c]), g1 = gpa_of(buf[c] + bytes / 2), g2 = gpa_of(buf[c] + bytes - 4096);
        group_rate(gpus, A, buf[c], slice, 1, nullptr);  // warm
        std::vector<double> pa, pb;
        const double ra = group_rate(gpus, A, buf[c], slice, 3, &pa);
        const double rb = group_rate(gpus, B, buf[c], slice, 3, &pb);
        std::printf("{\"chunk\": %d, \"mib\": %zu, \"alloc_s\": %.3f, \"gpa_first\": \"0x%llx\", \"gpa_mid\": \"0x%llx\", "
                    "\"gpa_last\": \"0x%llx\", \"group_a_gbs\": %.1f, \"group_b_gbs\": %.1f, \"a_min\": %.1f, \"b_min\": %.1f}\n",
                    c, mib, alloc_s, g0, g1, g2, ra, rb, *std::min_element(pa.begin(), pa.end()),
                    *std::min_element(pb.begin(), pb.end()));
        std::fflush(stdout);
    }
    return 0;
}
