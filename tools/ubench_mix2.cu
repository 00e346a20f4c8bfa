// ubench_mix2.cu — second pipe microbenchmark: where can the ACCUMULATE half of "compare + accumulate" run?
//
// The counting kernel issues one HSET2 (ALU pipe) and one HADD2 (fp16 FMA pipe) per two quartet x tree compares; that
// pair tops out at 3.2 warp-instr/clk/SM (tools/ubench_pipes.cu), i.e. 1.6 compares+accumulates.  This program times
// other accumulate forms next to an HSET2 with an integer-mask result (0xFFFF per true half):
//   IMAD   acc = mask * (-1) + acc        (integer multiply-add: FMA pipe; the multiplier comes from memory so that
//                                           ptxas cannot turn it into an IADD3)
//   IADD   acc = acc - mask               (whatever ptxas picks: IADD3 or IMAD.IADD)
//   VIADDMNMX (DPX) as the compare: relu(min(x - y, 1)) in {0,1} per int16 half, with IMAD / IADD accumulate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mix2 tools/ubench_mix2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int NCH = 16, ITERS = 4096;
enum Mix { M_HSET2_HADD2 = 0, M_HSETM_IMAD, M_HSETM_IADD, M_DPX, M_DPX_IMAD, M_DPX_IADD, M_HSET2, M_IMAD, M_2HSETM_IMAD_IADD3, MIX_COUNT };
static const char* mix_name[] = {"HSET2.BF + HADD2 (kernel today)", "HSET2(mask) + IMAD", "HSET2(mask) + IADD(2-input)", "VIADDMNMX alone", "VIADDMNMX + IMAD",
                                 "VIADDMNMX + IADD", "HSET2 alone", "IMAD alone", "4xHSET2(mask) + 2xIMAD + 1xIADD3(3in)"};
static const double instr_per_chain[] = {2, 2, 2, 1, 2, 2, 1, 1, 1.75};
static const double cmp_per_chain[] = {1, 1, 1, 1, 1, 1, 1, 0, 1};

#define HSET_BF(d, a, b) asm volatile("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define HSET_M(d, a, b) asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define HADD(acc, m) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(m))
#define IMAD(acc, m, k) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(m), "r"(k))
#define ISUB(acc, m) asm volatile("sub.s32 %0, %0, %1;" : "+r"(acc) : "r"(m))

template <int MIX>
__global__ void __launch_bounds__(512) k(uint32_t* out, const uint32_t* in) {
    uint32_t a[NCH], x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { a[i] = in[(threadIdx.x + i) & 255]; x[i] = in[(threadIdx.x * 3 + i * 7) & 255]; }
    const uint32_t mone = in[256], one2 = in[257];     // -1 and 0x00010001, unknown to the compiler
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            uint32_t m;
            const uint32_t other = a[(i + 5) % NCH];
            if (MIX == M_HSET2_HADD2) { HSET_BF(m, x[i], other); HADD(a[i], m); }
            if (MIX == M_HSETM_IMAD) { HSET_M(m, x[i], other); IMAD(a[i], m, mone); }
            if (MIX == M_HSETM_IADD) { HSET_M(m, x[i], other); ISUB(a[i], m); }
            if (MIX == M_DPX) { a[i] = __viaddmin_s16x2_relu(x[i], other, one2); }
            if (MIX == M_DPX_IMAD) { m = __viaddmin_s16x2_relu(x[i], other, one2); IMAD(a[i], m, mone); }
            if (MIX == M_DPX_IADD) { m = __viaddmin_s16x2_relu(x[i], other, one2); ISUB(a[i], m); }
            if (MIX == M_HSET2) { HSET_M(a[i], x[i], other); }
            if (MIX == M_IMAD) { IMAD(a[i], other, mone); }
            if (MIX == M_2HSETM_IMAD_IADD3) {
                // per 4 chains: 4 compares; two accumulate through IMAD (FMA pipe), the other two through one 3-input IADD3
                HSET_M(m, x[i], other);
                if ((i & 3) < 2) IMAD(a[i], m, mone);
                else if ((i & 3) == 2) { uint32_t m2; HSET_M(m2, x[i + 1], a[(i + 6) % NCH]); asm volatile("{.reg .s32 t; add.s32 t, %1, %2; sub.s32 %0, %0, t;}" : "+r"(a[i]) : "r"(m), "r"(m2)); }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
int run(int nsm, int threads, double clk_hz, uint32_t* out, uint32_t* in) {
    k<MIX><<<nsm, threads>>>(out, in);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MIX><<<nsm, threads>>>(out, in);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double chains = double(threads / 32) * ITERS * NCH;      // per SM
    const double clk = best * 1e-3 * clk_hz;
    printf("%-40s thr=%4d  time=%7.3f ms  winstr/clk/SM=%6.3f  compares(+acc)/clk/SM=%6.3f\n", mix_name[MIX], threads, best,
           chains * instr_per_chain[MIX] / clk, chains * cmp_per_chain[MIX] / clk);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max SM clock %d kHz (rates assume the kernels ran at this clock)\n", p.name, p.multiProcessorCount, clk_khz);
    uint32_t *out, *in;
    CK(cudaMalloc(&out, 4096 * 1024 * 4)); CK(cudaMalloc(&in, 2048));
    uint32_t h[258]; for (int i = 0; i < 256; i++) h[i] = 0x3c003c00u + (i * 0x00010001u);
    h[256] = 0xffffffffu; h[257] = 0x00010001u;
    CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
    for (int threads = 512; threads <= 1024; threads += 512) {
#define R(M) if (run<M>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in)) return 1
        R(M_HSET2_HADD2); R(M_HSETM_IMAD); R(M_HSETM_IADD); R(M_DPX); R(M_DPX_IMAD); R(M_DPX_IADD); R(M_HSET2); R(M_IMAD); R(M_2HSETM_IMAD_IADD3);
    }
    return 0;
}
