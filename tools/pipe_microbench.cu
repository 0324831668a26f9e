// pipe_microbench.cu -- issue-rate microbenchmark for the instruction classes
// the flagstat kernel is made of (LOP3, IMAD, SHF, HSET2, ...), alone and in
// the mixes the kernel uses.  Prints thread-instructions per clock per SM.
// The numbers feed the integer-pipe budget in DESIGN.md section 4.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

constexpr int kChains = 8;
constexpr int kInner = 16;  // ops per chain per loop iteration

enum Op {
    LOP3, IMAD, SHL, SHR, PRMT, IADD, POPC, HSET2, HFMA2, HADD2, IMADWIDE, FFMA,
    MIX_LOP3_IMAD, MIX_LOP3_HSET2, MIX_KERNEL, MIX_LOP3_SHR, MIX_LOP3_HFMA2, VOTE, SHFL,
    LDS, MIX_LOP3_LDS, MIX_LOP3_POPC, HMNMX2, MIX_LOP3_HSET2_IMAD,
    MINU16X2, ADDU16X2, IMNMX, FMNMX, SEL, MIX_LOP3_HMNMX2, MIX_LOP3_MINU16X2, MIX_LOP3_IADD, MIX_LOP3_FFMA,
    MIX_LOP3_HMNMX2_4, MIX_LOP3_HFMA2_8, MIX_LOP3_IADD_IMAD, HSET2F, MIX_LOP3_HFMA2_2, MIX_LOP3_HFMA2_3, MIX_LOP3_HFMA2_4, MIX_LOP3_HMUL2_3,
    MIX_LOP3_HFMA2_IMAD_6, NOPS
};

const char* kNames[] = {
    "LOP3", "IMAD", "SHF.L (funnel)", "SHF.R (funnel)", "PRMT", "IADD", "POPC", "HSET2 (set.eq.u32.f16x2)",
    "HFMA2", "HADD2", "IMAD.WIDE", "FFMA",
    "mix LOP3:IMAD 1:1", "mix LOP3:HSET2 1:1", "mix kernel 8 LOP3 : 2 HSET2 : 2 IMAD", "mix LOP3:SHR 1:1",
    "mix LOP3:HFMA2 1:1", "VOTE.ANY", "SHFL.BFLY", "LDS.32 (conflict-free)", "mix LOP3:LDS 4:1",
    "mix LOP3:POPC 4:1", "HMNMX2 (min.f16x2)", "mix LOP3:HSET2:IMAD 2:1:1",
    "min.u16x2", "add.u16x2", "min.u32", "min.f32", "selp", "mix LOP3:HMNMX2 1:1", "mix LOP3:min.u16x2 1:1",
    "mix LOP3:IADD 1:1", "mix LOP3:FFMA 1:1", "mix LOP3:HMNMX2 4:1", "mix LOP3:HFMA2 8:1", "mix LOP3:IADD:IMAD 2:1:1",
    "HSET2 (set.eq.f16x2.f16x2)", "mix LOP3:HFMA2 2:1", "mix LOP3:HFMA2 3:1", "mix LOP3:HFMA2 4:1", "mix LOP3:HMUL2 3:1",
    "mix LOP3:HFMA2:IMAD 6:2:1"};

template <int OP>
__device__ __forceinline__ void body(uint32_t (&a)[kChains], uint32_t b, uint32_t c, const uint32_t* sm)
{
#pragma unroll
    for (int k = 0; k < kInner; ++k) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            const uint32_t y = a[(i + 1) % kChains];
            (void)y;
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
            if (OP == SHL) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(y));
            if (OP == SHR) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(y));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
            if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            if (OP == HSET2) asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == IMADWIDE) {
                uint64_t w;
                asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(a[i]), "r"(b), "l"((uint64_t)c));
                a[i] = (uint32_t)(w >> 32) ^ (uint32_t)w;
            }
            if (OP == FFMA) {
                float f = __uint_as_float(a[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
                a[i] = __float_as_uint(f);
            }
            if (OP == MIX_LOP3_IMAD) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_HSET2) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            }
            if (OP == MIX_LOP3_HFMA2) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (OP == MIX_LOP3_SHR) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(y));
            }
            if (OP == MIX_KERNEL) {  // per 12 ops: 8 LOP3, 2 HSET2, 2 IMAD  (kInner = 16 -> pattern of 4: L L L X)
                const int m = (k * kChains + i) % 6;
                if (m == 4) asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
                else if (m == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_HSET2_IMAD) {
                const int m = k % 4;
                if (m == 2) asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
                else if (m == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }

            if (OP == MINU16X2) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == ADDU16X2) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == IMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == FMNMX) {
                float f = __uint_as_float(a[i]);
                asm volatile("min.f32 %0, %0, %1;" : "+f"(f) : "f"(__uint_as_float(y)));
                a[i] = __float_as_uint(f);
            }
            if (OP == SEL) asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; selp.u32 %0, %0, %1, q; }" : "+r"(a[i]) : "r"(y), "r"(c));
            if (OP == HSET2F) asm volatile("set.eq.f16x2.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            if (OP == MIX_LOP3_HMNMX2) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            }
            if (OP == MIX_LOP3_MINU16X2) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            }
            if (OP == MIX_LOP3_IADD) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
            }
            if (OP == MIX_LOP3_FFMA) {
                if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
                else {
                    float f = __uint_as_float(a[i]);
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
                    a[i] = __float_as_uint(f);
                }
            }
            if (OP == MIX_LOP3_HMNMX2_4) {
                if (k % 5 == 4) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_HFMA2_8) {
                if ((k * kChains + i) % 9 == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_IADD_IMAD) {
                const int m = k % 4;
                if (m == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y));
                else if (m == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }

            if (OP == MIX_LOP3_HFMA2_2 || OP == MIX_LOP3_HFMA2_3 || OP == MIX_LOP3_HFMA2_4) {
                const int per = OP == MIX_LOP3_HFMA2_2 ? 3 : OP == MIX_LOP3_HFMA2_3 ? 4 : 5;
                if ((k * kChains + i) % per == per - 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_HMUL2_3) {
                if ((k * kChains + i) % 4 == 3) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_HFMA2_IMAD_6) {
                const int m = (k * kChains + i) % 9;
                if (m == 2 || m == 6) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                else if (m == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == VOTE) {
                uint32_t p;
                asm volatile("{ .reg .pred q, r; setp.ne.u32 q, %1, 0; vote.sync.any.pred r, q, 0xffffffff; selp.u32 %0, 1, 0, r; }"
                             : "=r"(p) : "r"(a[i]));
                a[i] += p;
            }
            if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(a[i]));
            if (OP == LDS) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm + ((a[i] + threadIdx.x) & 1023))));
                a[i] = v;
            }
            if (OP == MIX_LOP3_LDS) {
                if (k % 5 == 4) {
                    uint32_t v;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm + ((a[i] + threadIdx.x) & 1023))));
                    a[i] = v;
                } else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
            if (OP == MIX_LOP3_POPC) {
                if (k % 5 == 4) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(y), "r"(c));
            }
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, long long* cycles, int iters, uint32_t b, uint32_t c)
{
    __shared__ uint32_t sm[1024];
    sm[threadIdx.x] = (threadIdx.x * 7u) & 1023u;
    __syncthreads();
    uint32_t a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = threadIdx.x * 2654435761u + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) body<OP>(a, b, c, sm);
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) x ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(uint32_t* d_out, long long* d_cyc, int nsm)
{
    const int iters = 200;
    bench<OP><<<nsm, 1024>>>(d_out, d_cyc, 10, 3u, 5u);  // warm-up
    bench<OP><<<nsm, 1024>>>(d_out, d_cyc, iters, 3u, 5u);
    CHECK(cudaDeviceSynchronize());
    long long cyc[256];
    CHECK(cudaMemcpy(cyc, d_cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += (double)cyc[i];
    avg /= nsm;
    const double ops = (double)iters * kInner * kChains * 1024.0;
    std::printf("%-40s %8.2f thread-instr/clk/SM\n", kNames[OP], ops / avg);
    return 0;
}

int main()
{
    int nsm = 0;
    CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    std::printf("device: %s, %d SMs\n", prop.name, nsm);
    uint32_t* d_out;
    long long* d_cyc;
    CHECK(cudaMalloc(&d_out, sizeof(uint32_t) * 1024 * nsm));
    CHECK(cudaMalloc(&d_cyc, sizeof(long long) * nsm));
    int rc = 0;
    rc |= run<LOP3>(d_out, d_cyc, nsm);
    rc |= run<IMAD>(d_out, d_cyc, nsm);
    rc |= run<SHL>(d_out, d_cyc, nsm);
    rc |= run<SHR>(d_out, d_cyc, nsm);
    rc |= run<PRMT>(d_out, d_cyc, nsm);
    rc |= run<IADD>(d_out, d_cyc, nsm);
    rc |= run<POPC>(d_out, d_cyc, nsm);
    rc |= run<HSET2>(d_out, d_cyc, nsm);
    rc |= run<HFMA2>(d_out, d_cyc, nsm);
    rc |= run<HADD2>(d_out, d_cyc, nsm);
    rc |= run<HMNMX2>(d_out, d_cyc, nsm);
    rc |= run<IMADWIDE>(d_out, d_cyc, nsm);
    rc |= run<FFMA>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_IMAD>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HSET2>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_SHR>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HSET2_IMAD>(d_out, d_cyc, nsm);
    rc |= run<MIX_KERNEL>(d_out, d_cyc, nsm);
    rc |= run<VOTE>(d_out, d_cyc, nsm);
    rc |= run<SHFL>(d_out, d_cyc, nsm);
    rc |= run<LDS>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_LDS>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_POPC>(d_out, d_cyc, nsm);
    rc |= run<MINU16X2>(d_out, d_cyc, nsm);
    rc |= run<ADDU16X2>(d_out, d_cyc, nsm);
    rc |= run<IMNMX>(d_out, d_cyc, nsm);
    rc |= run<FMNMX>(d_out, d_cyc, nsm);
    rc |= run<SEL>(d_out, d_cyc, nsm);
    rc |= run<HSET2F>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HMNMX2>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_MINU16X2>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_IADD>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_FFMA>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HMNMX2_4>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2_8>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_IADD_IMAD>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2_2>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2_3>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2_4>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HMUL2_3>(d_out, d_cyc, nsm);
    rc |= run<MIX_LOP3_HFMA2_IMAD_6>(d_out, d_cyc, nsm);
    return rc;
}
