// Microbenchmark: issue throughput of the instruction mixes the counting kernel could be built from.
// Reports warp-instructions / clk / SM for each mix from the kernel's wall time (CUDA events) at the
// SM clock the device reports under load, so the counting kernel's design is chosen from
// measurements on the B200, not from guesses.  Not part of the product path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
//
// Every compare reads an accumulator of ANOTHER chain, so ptxas cannot hoist it out of the loop
// (round-1a's version compared two loop-invariant registers and ptxas hoisted all HSET2: its
// "HSET2+HADD2 = 4/clk/SM" line measured HADD2 alone).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NCH = 16;      // independent chains per thread
constexpr int ITERS = 16384; // loop iterations

enum Mix { M_IADD3 = 0, M_LOP3, M_IMAD, M_HADD2, M_HFMA2, M_HSET2, M_HSET2_HADD2, M_HSET2_HFMA2, M_HSET2_ALT, M_HSAT_HADD2,
           M_2HSET2_IADD3, M_CMPBOTH_HADD2, M_HMMA, M_4HSET2_HMMA, M_CMPBOTH_HMMA, M_MIX_TC, M_REAL_LDS, M_REAL_NOLDS, MIX_COUNT };
static const char* mix_name[] = {"IADD3", "LOP3", "IMAD", "HADD2", "HFMA2", "HSET2", "HSET2+HADD2", "HSET2+HFMA2", "HSET2+{HADD2|HFMA2}",
    "HADD2.SAT(cmp)+HADD2", "2xHSET2(mask)+IADD3(3in)", "{HSET2|HADD2.SAT}+HADD2", "HMMA.16816.F16 alone", "4xHSET2+HMMA", "2xHSET2+2xHADD2.SAT+HMMA",
    "8cmp(4A+4F)+4HADD2+1HMMA", "count loop model (8 LDS.128/192)", "count loop model (no LDS)"};
// instructions per chain per iteration (for the rate computation)
static const double mix_ipc[] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 1.5, 2, 0.25, 1.25, 1.25, 13.0 / 8.0, 2.0 + 16.0 / 96.0 + 8.0 / 96.0, 2.0 + 4.0 / 96.0};

__device__ __forceinline__ void hmma(uint32_t& d0, uint32_t& d1, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                 : "+r"(d0), "+r"(d1) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
#define HSET_GT(m, p, q) asm volatile("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(m) : "r"(p), "r"(q))
#define HSET_LT(m, p, q) asm volatile("set.lt.f16x2.f16x2 %0, %1, %2;" : "=r"(m) : "r"(p), "r"(q))
#define HSETM_GT(m, p, q) asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(p), "r"(q))
#define HSAT(m, p, q) asm volatile("sub.sat.f16x2 %0, %1, %2;" : "=r"(m) : "r"(p), "r"(q))
#define HADD(acc, m) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(m))
#define HFMA1(acc, m, one) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(acc) : "r"(m), "r"(one))

template <int MIX>
__global__ void __launch_bounds__(1024) k(uint32_t* out, const uint32_t* in) {
    uint32_t a[NCH], x[NCH];
    __shared__ uint4 sm[256];
    if (threadIdx.x < 256) sm[threadIdx.x] = make_uint4(in[threadIdx.x & 63], in[(threadIdx.x + 1) & 63], in[(threadIdx.x + 2) & 63], in[(threadIdx.x + 3) & 63]);
    __syncthreads();
    uint32_t y = in[threadIdx.x & 31], z = in[(threadIdx.x + 7) & 31];
    const uint32_t one = 0x3c003c00u;
#pragma unroll
    for (int i = 0; i < NCH; i++) { a[i] = in[(threadIdx.x + i) & 63]; x[i] = in[(threadIdx.x * 3 + i) & 63]; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MIX == M_REAL_LDS || MIX == M_REAL_NOLDS) {
            // one tree step of the counting kernel: 8 LDS.128, 16 HSUB2, 96 HSET2 (operand selectors), 96 HADD2 on 96 accumulators
            // modelled with NCH=16 accumulators reused 6x (dependency distance 16, as in the real loop the distance is 96)
            uint4 v0, v1;
            if (MIX == M_REAL_LDS) { v0 = sm[(it + threadIdx.x) & 255]; v1 = sm[(it * 3 + threadIdx.x) & 255]; }
            else { v0 = make_uint4(x[0], x[1], x[2], x[3]); v1 = make_uint4(x[4], x[5], x[6], x[7]); }
            uint32_t g[4];
            asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[0]) : "r"(v0.x), "r"(v1.x));
            asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[1]) : "r"(v0.y), "r"(v1.y));
            asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[2]) : "r"(v0.z), "r"(v1.z));
            asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[3]) : "r"(v0.w), "r"(v1.w));
#pragma unroll
            for (int r = 0; r < 6; r++) {
#pragma unroll
                for (int i = 0; i < NCH; i++) {
                    uint32_t m;
                    if (r & 1) HSET_GT(m, g[i & 3], a[(i + 5) % NCH]); else HSET_LT(m, g[(i + 1) & 3], a[(i + 9) % NCH]);
                    HADD(a[i], m);
                }
                if (MIX == M_REAL_LDS && r < 3) {   // the remaining 6 LDS.128 (8 per tree step, 2 issued above)
                    uint4 w0 = sm[(it + r * 7 + threadIdx.x) & 255], w1 = sm[(it * 5 + r + threadIdx.x) & 255];
                    asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[0]) : "r"(w0.x), "r"(w1.x));
                    asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[1]) : "r"(w0.y), "r"(w1.y));
                    asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[2]) : "r"(w0.z), "r"(w1.z));
                    asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(g[3]) : "r"(w0.w), "r"(w1.w));
                }
            }
            continue;
        }
        if (MIX == M_HMMA) {
#pragma unroll
            for (int i = 0; i < NCH; i += 4) hmma(a[i], a[i + 1], x[i], x[i + 1], x[i + 2], x[i + 3], y, z);
            continue;
        }
        if (MIX == M_4HSET2_HMMA || MIX == M_CMPBOTH_HMMA) {
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                uint32_t m0, m1, m2, m3;
                HSET_GT(m0, x[i], a[(i + 5) % NCH]);
                HSET_GT(m1, x[i + 1], a[(i + 6) % NCH]);
                if (MIX == M_4HSET2_HMMA) { HSET_GT(m2, x[i + 2], a[(i + 7) % NCH]); HSET_GT(m3, x[i + 3], a[(i + 8) % NCH]); }
                else { HSAT(m2, x[i + 2], a[(i + 7) % NCH]); HSAT(m3, x[i + 3], a[(i + 8) % NCH]); }
                hmma(a[i], a[i + 1], m0, m1, m2, m3, one, one);
            }
            continue;
        }
        if (MIX == M_MIX_TC) {
            // per 8 compares: 4 on ALU (HSET2), 4 on the fp16 pipe (HADD2.SAT); 4 accumulated by HADD2, 4 by one HMMA
#pragma unroll
            for (int i = 0; i < NCH; i += 8) {
                uint32_t m0, m1, m2, m3, m4, m5, m6, m7;
                HSET_GT(m0, x[i], a[(i + 9) % NCH]); HSAT(m1, x[i + 1], a[(i + 10) % NCH]);
                HSET_GT(m2, x[i + 2], a[(i + 11) % NCH]); HSAT(m3, x[i + 3], a[(i + 12) % NCH]);
                HSET_GT(m4, x[i + 4], a[(i + 13) % NCH]); HSAT(m5, x[i + 5], a[(i + 14) % NCH]);
                HSET_GT(m6, x[i + 6], a[(i + 15) % NCH]); HSAT(m7, x[i + 7], a[(i + 8) % NCH]);
                hmma(a[i], a[i + 1], m0, m1, m2, m3, one, one);
                HADD(a[i + 2], m4); HADD(a[i + 3], m5); HADD(a[i + 4], m6); HADD(a[i + 5], m7);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            uint32_t m;
            if (MIX == M_IADD3) { asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 5) % NCH])); }
            if (MIX == M_LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[(i + 5) % NCH]), "r"(y)); }
            if (MIX == M_IMAD) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(a[(i + 5) % NCH]), "r"(y)); }
            if (MIX == M_HADD2) { HADD(a[i], a[(i + 5) % NCH]); }
            if (MIX == M_HFMA2) { HFMA1(a[i], a[(i + 5) % NCH], one); }
            if (MIX == M_HSET2) { HSET_GT(a[i], x[i], a[(i + 5) % NCH]); }
            if (MIX == M_HSET2_HADD2) { HSET_GT(m, x[i], a[(i + 5) % NCH]); HADD(a[i], m); }
            if (MIX == M_HSET2_HFMA2) { HSET_GT(m, x[i], a[(i + 5) % NCH]); HFMA1(a[i], m, one); }
            if (MIX == M_HSET2_ALT) { HSET_GT(m, x[i], a[(i + 5) % NCH]); if (i & 1) HADD(a[i], m); else HFMA1(a[i], m, one); }
            if (MIX == M_HSAT_HADD2) { HSAT(m, x[i], a[(i + 5) % NCH]); HADD(a[i], m); }
            if (MIX == M_CMPBOTH_HADD2) { if (i & 1) HSET_GT(m, x[i], a[(i + 5) % NCH]); else HSAT(m, x[i], a[(i + 5) % NCH]); HADD(a[i], m); }
            if (MIX == M_2HSET2_IADD3) {
                if (i & 1) {
                    uint32_t m2;
                    HSETM_GT(m, x[i], a[(i + 5) % NCH]); HSETM_GT(m2, x[i - 1], a[(i + 6) % NCH]);
                    asm volatile("{.reg .s32 t; add.s32 t, %1, %2; sub.s32 %0, %0, t;}" : "+r"(a[i]) : "r"(m), "r"(m2));   // ptxas -> IADD3 a, a, -m, -m2
                }
            }
        }
    }
    uint32_t s = y ^ z;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= a[i] ^ x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
void run(int nsm, int threads, int ctas_per_sm, double clk_hz, uint32_t* out, uint32_t* in) {
    int grid = nsm * ctas_per_sm;
    k<MIX><<<grid, threads>>>(out, in);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MIX><<<grid, threads>>>(out, in);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double warps_per_sm = double(threads / 32) * ctas_per_sm;
    double winstr_per_sm = warps_per_sm * ITERS * NCH * mix_ipc[MIX];
    if (MIX == M_REAL_LDS || MIX == M_REAL_NOLDS) winstr_per_sm = warps_per_sm * ITERS * 96.0 * mix_ipc[MIX];
    double rate = winstr_per_sm / (best * 1e-3 * clk_hz);   // warp-instr / clk / SM   (4.0 = one per SMSP per clk)
    double pairs = (MIX == M_REAL_LDS || MIX == M_REAL_NOLDS) ? warps_per_sm * ITERS * 96.0 : 0;   // compare+accumulate pairs per SM
    printf("%-36s thr=%4d  time=%8.3f ms  winstr/clk/SM=%6.3f", mix_name[MIX], threads, best, rate);
    if (pairs > 0) printf("  cmp+acc pairs/clk/SM=%6.3f (2.0 = both half-rate pipes saturated)", pairs / (best * 1e-3 * clk_hz));
    printf("\n");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clk_hz = clk_khz * 1e3;
    printf("device %s, %d SMs, max SM clock %d kHz (rates assume the kernel ran at this clock)\n", p.name, nsm, clk_khz);
    uint32_t *out, *in;
    CK(cudaMalloc(&out, 4096 * 1024 * 4)); CK(cudaMalloc(&in, 1024));
    uint32_t h[256]; for (int i = 0; i < 256; i++) h[i] = 0x3c003c00u + (i * 0x00010001u);  // small halfs ~1.0..
    CK(cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice));
    for (int cfg = 0; cfg < 2; cfg++) {
        int threads = cfg == 0 ? 512 : 1024;
        printf("--- %d threads/SM ---\n", threads);
#define R(M) run<M>(nsm, threads, 1, clk_hz, out, in)
        R(M_IADD3); R(M_LOP3); R(M_IMAD); R(M_HADD2); R(M_HFMA2); R(M_HSET2); R(M_HSET2_HADD2); R(M_HSET2_HFMA2); R(M_HSET2_ALT);
        R(M_HSAT_HADD2); R(M_2HSET2_IADD3); R(M_CMPBOTH_HADD2); R(M_HMMA); R(M_4HSET2_HMMA); R(M_CMPBOTH_HMMA); R(M_MIX_TC);
        R(M_REAL_LDS); R(M_REAL_NOLDS);
    }
    return 0;
}
