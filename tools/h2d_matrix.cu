// h2d_matrix.cu -- where does host->device bandwidth go when several GPUs of one box pull
// pinned host memory at the same time?  (e2e per-GPU H2D rate fell
// from 55 GB/s at 1-2 GPUs to 28.8 at 4 and 23.3 at 8.)
//
// One process, one thread per GPU.  For every subset in {each GPU alone, every pair, a few
// groups of 4, all} the chosen GPUs copy B bytes of pinned host memory R times with
// cudaMemcpyAsync on their own streams, all started behind one barrier; per-GPU rates come
// from CUDA events, the aggregate from the wall clock around the whole subset.  With more than
// one NUMA node visible the singles are repeated with the pinned buffer bound to each node
// (mbind + cudaHostRegister) and the issuing thread pinned to that node's CPUs.
// The PCIe path of every GPU (sysfs: bridges above it, link speed / width, numa_node) is
// printed first, so that GPUs sharing an upstream switch port can be told from the matrix.
//
//   nvcc -O2 -std=c++17 -o tools/bin/h2d_matrix tools/h2d_matrix.cu -lpthread
//   tools/bin/h2d_matrix [MiB per copy = 512] [repeats = 6]     -> JSON lines on stdout
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
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

static std::string slurp(const std::string& path)
{
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    return s;
}

static std::string json_escape(const std::string& s)
{
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        if (c == '\n') { o += "\\n"; continue; }
        o += c;
    }
    return o;
}

// "0-3,8,10-11" -> cpu list
static std::vector<int> parse_cpulist(const std::string& s)
{
    std::vector<int> v;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        const size_t dash = tok.find('-');
        const int a = std::atoi(tok.c_str());
        const int b = dash == std::string::npos ? a : std::atoi(tok.c_str() + dash + 1);
        for (int c = a; c <= b; ++c) v.push_back(c);
    }
    return v;
}

struct Node {
    int id;
    std::vector<int> cpus;
};

static std::vector<Node> numa_nodes()
{
    std::vector<Node> nodes;
    for (int n = 0; n < 64; ++n) {
        const std::string base = "/sys/devices/system/node/node" + std::to_string(n);
        std::ifstream f(base + "/cpulist");
        if (!f.good()) continue;
        nodes.push_back(Node{n, parse_cpulist(slurp(base + "/cpulist"))});
    }
    return nodes;
}

static void pin_thread(const std::vector<int>& cpus)
{
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

// page-aligned anonymous memory bound to one NUMA node (MPOL_BIND = 2), touched, then registered
static void* alloc_on_node(size_t bytes, int node, bool* bound)
{
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    *bound = false;
    if (node >= 0) {
        unsigned long mask[16] = {0};
        mask[node / 64] |= 1ul << (node % 64);
        const long rc = syscall(SYS_mbind, p, bytes, 2 /*MPOL_BIND*/, mask, 1024ul, 0u);
        *bound = rc == 0;
    }
    std::memset(p, 0x5a, bytes);
    return p;
}

struct Gpu {
    int dev;
    char busid[32];
    void* d = nullptr;
    void* h = nullptr;  // default pinned buffer (cudaHostAlloc)
    cudaStream_t st;
    cudaEvent_t e0, e1;
};

struct Barrier {
    std::atomic<int> count{0};
    std::atomic<int> gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) {
            count.store(0);
            gen.fetch_add(1);
        } else {
            while (gen.load() == g) std::this_thread::yield();
        }
    }
};

// copies from hsrc[i] (or the GPU's own pinned buffer) on every GPU of `sel`, concurrently
static void run_subset(std::vector<Gpu>& gpus, const std::vector<int>& sel, size_t bytes, int reps,
                       const char* label, const std::vector<void*>* hsrc = nullptr,
                       const std::vector<std::vector<int>>* cpus = nullptr, const char* extra = "")
{
    const int k = (int)sel.size();
    Barrier bar(k + 1);
    std::vector<float> ms(k, 0.f);
    std::vector<std::thread> th;
    for (int i = 0; i < k; ++i) {
        th.emplace_back([&, i] {
            Gpu& g = gpus[sel[i]];
            if (cpus) pin_thread((*cpus)[i]);
            CK(cudaSetDevice(g.dev));
            const void* src = hsrc ? (*hsrc)[i] : g.h;
            CK(cudaMemcpyAsync(g.d, src, bytes, cudaMemcpyHostToDevice, g.st));  // warm
            CK(cudaStreamSynchronize(g.st));
            bar.wait();
            CK(cudaEventRecord(g.e0, g.st));
            for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(g.d, src, bytes, cudaMemcpyHostToDevice, g.st));
            CK(cudaEventRecord(g.e1, g.st));
            CK(cudaStreamSynchronize(g.st));
            bar.wait();
            CK(cudaEventElapsedTime(&ms[i], g.e0, g.e1));
        });
    }
    bar.wait();
    const auto t0 = std::chrono::steady_clock::now();
    bar.wait();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& t : th) t.join();
    std::printf("{\"test\": \"%s\", \"gpus\": [", label);
    for (int i = 0; i < k; ++i) std::printf("%s%d", i ? ", " : "", gpus[sel[i]].dev);
    std::printf("], \"gbs_per_gpu\": [");
    double sum = 0;
    for (int i = 0; i < k; ++i) {
        const double gbs = (double)bytes * reps / (ms[i] * 1e-3) / 1e9;
        sum += gbs;
        std::printf("%s%.2f", i ? ", " : "", gbs);
    }
    std::printf("], \"sum_of_event_rates_gbs\": %.2f, \"aggregate_wall_gbs\": %.2f%s}\n", sum,
                (double)bytes * reps * k / wall / 1e9, extra);
    std::fflush(stdout);
}

int main(int argc, char** argv)
{
    const size_t mib = argc > 1 ? (size_t)std::atol(argv[1]) : 512;
    const int reps = argc > 2 ? std::atoi(argv[2]) : 6;
    const size_t bytes = mib << 20;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const std::vector<Node> nodes = numa_nodes();

    // ---- host / PCIe description ----
    std::printf("{\"host\": {\"numa_nodes\": %zu, \"online_cpus\": \"%s\", \"nodes\": [", nodes.size(),
                json_escape(slurp("/sys/devices/system/cpu/online")).c_str());
    for (size_t i = 0; i < nodes.size(); ++i)
        std::printf("%s{\"id\": %d, \"cpus\": %zu, \"meminfo_first_line\": \"%s\"}", i ? ", " : "", nodes[i].id,
                    nodes[i].cpus.size(),
                    json_escape(slurp("/sys/devices/system/node/node" + std::to_string(nodes[i].id) + "/meminfo")
                                    .substr(0, 60)).c_str());
    std::printf("]}, \"mib_per_copy\": %zu, \"repeats\": %d}\n", mib, reps);

    std::vector<Gpu> gpus(ndev);
    for (int g = 0; g < ndev; ++g) {
        Gpu& G = gpus[g];
        G.dev = g;
        CK(cudaSetDevice(g));
        CK(cudaDeviceGetPCIBusId(G.busid, sizeof(G.busid), g));
        for (char* c = G.busid; *c; ++c) *c = (char)std::tolower(*c);
        CK(cudaMalloc(&G.d, bytes));
        CK(cudaHostAlloc(&G.h, bytes, cudaHostAllocPortable));
        std::memset(G.h, 0x33, bytes);
        CK(cudaStreamCreateWithFlags(&G.st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&G.e0));
        CK(cudaEventCreate(&G.e1));
        const std::string sys = std::string("/sys/bus/pci/devices/") + G.busid;
        char real[4096] = {0};
        if (!realpath(sys.c_str(), real)) real[0] = 0;
        std::printf("{\"gpu\": %d, \"busid\": \"%s\", \"sysfs_path\": \"%s\", \"numa_node\": \"%s\", "
                    "\"link_speed\": \"%s\", \"link_width\": \"%s\", \"max_link_speed\": \"%s\", \"max_link_width\": \"%s\"}\n",
                    g, G.busid, json_escape(real).c_str(), slurp(sys + "/numa_node").c_str(),
                    slurp(sys + "/current_link_speed").c_str(), slurp(sys + "/current_link_width").c_str(),
                    slurp(sys + "/max_link_speed").c_str(), slurp(sys + "/max_link_width").c_str());
    }
    std::fflush(stdout);

    // ---- singles ----
    for (int g = 0; g < ndev; ++g) run_subset(gpus, {g}, bytes, reps, "single");
    // ---- every pair ----
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b) run_subset(gpus, {a, b}, bytes, reps, "pair");
    // ---- groups ----
    if (ndev >= 4) {
        run_subset(gpus, {0, 1, 2, 3}, bytes, reps, "four_0123");
        if (ndev >= 8) {
            run_subset(gpus, {4, 5, 6, 7}, bytes, reps, "four_4567");
            run_subset(gpus, {0, 2, 4, 6}, bytes, reps, "four_0246");
            run_subset(gpus, {1, 3, 5, 7}, bytes, reps, "four_1357");
            run_subset(gpus, {0, 1, 4, 5}, bytes, reps, "four_0145");
            run_subset(gpus, {0, 3, 4, 7}, bytes, reps, "four_0347");
        }
    }
    if (ndev >= 2) {
        std::vector<int> all;
        for (int g = 0; g < ndev; ++g) all.push_back(g);
        run_subset(gpus, all, bytes, reps, "all");
        // longer run of everything together (steady state rather than a burst)
        run_subset(gpus, all, bytes, reps * 4, "all_long");
    }

    // ---- NUMA placement: buffer bound to node n, issuing thread pinned to node c ----
    if (nodes.size() > 1) {
        for (const Node& nd : nodes) {
            bool bound = false;
            void* p = alloc_on_node(bytes, nd.id, &bound);
            if (!p) continue;
            if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
                cudaGetLastError();
                munmap(p, bytes);
                continue;
            }
            for (int g = 0; g < ndev; ++g)
                for (const Node& cn : nodes) {
                    std::vector<void*> src{p};
                    std::vector<std::vector<int>> cp{cn.cpus};
                    char extra[128];
                    std::snprintf(extra, sizeof extra, ", \"buffer_node\": %d, \"mbind_ok\": %s, \"thread_node\": %d",
                                  nd.id, bound ? "true" : "false", cn.id);
                    run_subset(gpus, {g}, bytes, reps, "single_numa", &src, &cp, extra);
                }
            cudaHostUnregister(p);
            munmap(p, bytes);
        }
        // all GPUs together, every GPU's buffer on its own sysfs numa_node (or round-robin if unknown)
        std::vector<void*> src;
        std::vector<std::vector<int>> cp;
        std::vector<int> all;
        std::string placement;
        for (int g = 0; g < ndev; ++g) {
            int want = std::atoi(slurp(std::string("/sys/bus/pci/devices/") + gpus[g].busid + "/numa_node").c_str());
            if (want < 0 || want >= (int)nodes.size()) want = nodes[(size_t)g * nodes.size() / ndev].id;
            bool bound = false;
            void* p = alloc_on_node(bytes, want, &bound);
            if (!p || cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
                std::fprintf(stderr, "numa placement: allocation for gpu %d failed\n", g);
                return 0;
            }
            src.push_back(p);
            for (const Node& nd : nodes)
                if (nd.id == want) cp.push_back(nd.cpus);
            if ((int)cp.size() < g + 1) cp.push_back({});
            all.push_back(g);
            placement += std::to_string(want) + (g + 1 < ndev ? "," : "");
        }
        const std::string extra = ", \"buffer_nodes\": \"" + placement + "\"";
        run_subset(gpus, all, bytes, reps * 2, "all_numa_local", &src, &cp, extra.c_str());
    } else {
        std::printf("{\"note\": \"one NUMA node visible: placement cannot be controlled from inside this VM\"}\n");
    }
    return 0;
}
