// Microbenchmark 2: the counting kernel's real inner step (kernels/count_roles.cuh) fed from shared memory,
// in the variants considered for the data layout.  Reports compares (HSET2 warp-instr) per clk per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_loop ubench_loop.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../quartetscores_b200/csrc/kernels/count_roles.cuh"
using namespace qs;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
enum { V_SEL_IADD3 = 0, V_TP_IADD3, V_SEL_HADD2, V_SEL_IADD3_NOLDS, V_TP_IADD3_NOLDS, V_PRMT_IADD3, NV };
static const char* vname[] = {"a-packed, H0_H0 selectors, HSET2 mask + IADD3 (round-1c kernel)", "tree-pair packed (no selectors), HSET2 mask + IADD3",
                              "a-packed, selectors, HSET2.BF + HADD2 (round-1a kernel)", "a-packed selectors + IADD3, operands from registers (no LDS)",
                              "tree-pair packed + IADD3, operands from registers (no LDS)", "a-packed, explicit PRMT broadcast + IADD3"};

// tree-pair packed step: 8 u x 4 v quartets, each register = (tree 2k, tree 2k+1); two tree pairs per call
struct TPCounters { uint32_t gt[4][8], lt[4][8]; };
__device__ __forceinline__ void step_tp(TPCounters& c, const uint4 (&pu0)[2], const uint4 (&qu0)[2], const uint4& pv0, const uint4& qv0,
                                        const uint4 (&pu1)[2], const uint4 (&qu1)[2], const uint4& pv1, const uint4& qv1) {
    __half2 u0[8], u1[8], v0[4], v1[4];
    sub4(*reinterpret_cast<__half2(*)[4]>(&u0[0]), qu0[0], pu0[0]); sub4(*reinterpret_cast<__half2(*)[4]>(&u0[4]), qu0[1], pu0[1]);
    sub4(*reinterpret_cast<__half2(*)[4]>(&u1[0]), qu1[0], pu1[0]); sub4(*reinterpret_cast<__half2(*)[4]>(&u1[4]), qu1[1], pu1[1]);
    sub4(v0, qv0, pv0); sub4(v1, qv1, pv1);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            c.gt[j][p] = c.gt[j][p] - __hgt2_mask(u0[p], v0[j]) - __hgt2_mask(u1[p], v1[j]);
            c.lt[j][p] = c.lt[j][p] - __hlt2_mask(u0[p], v0[j]) - __hlt2_mask(u1[p], v1[j]);
        }
}

struct HCounters { __half2 gt[8][4], lt[8][4]; };
__device__ __forceinline__ void step_hadd(HCounters& x, const BlockRows& r0, const BlockRows& r1) {
    __half2 u0[4], v0[4], u1[4], v1[4];
    sub4(u0, r0.qu, r0.pu); sub4(v0, r0.qv, r0.pv); sub4(u1, r1.qu, r1.pu); sub4(v1, r1.qv, r1.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b0 = (j & 1) ? __high2half2(v0[j >> 1]) : __low2half2(v0[j >> 1]);
        const __half2 b1 = (j & 1) ? __high2half2(v1[j >> 1]) : __low2half2(v1[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            x.gt[j][p] = __hadd2(__hadd2(x.gt[j][p], __hgt2(u0[p], b0)), __hgt2(u1[p], b1));
            x.lt[j][p] = __hadd2(__hadd2(x.lt[j][p], __hlt2(u0[p], b0)), __hlt2(u1[p], b1));
        }
    }
}
__device__ __forceinline__ void step_prmt(XCounters& x, const BlockRows& r0, const BlockRows& r1) {
    __half2 u0[4], v0[4], u1[4], v1[4];
    sub4(u0, r0.qu, r0.pu); sub4(v0, r0.qv, r0.pv); sub4(u1, r1.qu, r1.pu); sub4(v1, r1.qv, r1.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t w0 = *reinterpret_cast<uint32_t*>(&v0[j >> 1]), w1 = *reinterpret_cast<uint32_t*>(&v1[j >> 1]), b0, b1;
        asm volatile("prmt.b32 %0, %1, %1, %2;" : "=r"(b0) : "r"(w0), "r"((j & 1) ? 0x3232u : 0x1010u));
        asm volatile("prmt.b32 %0, %1, %1, %2;" : "=r"(b1) : "r"(w1), "r"((j & 1) ? 0x3232u : 0x1010u));
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            x.gt[j][p] = x.gt[j][p] - __hgt2_mask(u0[p], as_h2(b0)) - __hgt2_mask(u1[p], as_h2(b1));
            x.lt[j][p] = x.lt[j][p] - __hlt2_mask(u0[p], as_h2(b0)) - __hlt2_mask(u1[p], as_h2(b1));
        }
    }
}

template <int V>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, const uint32_t* in) {
    extern __shared__ __align__(16) unsigned char sm[];     // 64 KB of "matrix rows"
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = in[i & 255];
    __syncthreads();
    const uint32_t o0 = (threadIdx.x * 16u) & 0x3ff0u, o1 = (threadIdx.x * 48u + 4096u) & 0x3ff0u, o2 = (threadIdx.x * 80u + 8192u) & 0x3ff0u, o3 = ((threadIdx.x >> 3) * 16u + 12288u) & 0x3ff0u;
    uint32_t acc = 0;
    if (V == V_SEL_IADD3 || V == V_SEL_IADD3_NOLDS || V == V_PRMT_IADD3) {
        XCounters x; zero(x);
        BlockRows r0, r1;
        r0 = BlockRows{lds128(sm, o0), lds128(sm, o1), lds128(sm, o2), lds128(sm, o3)}; r1 = BlockRows{lds128(sm, o1), lds128(sm, o2), lds128(sm, o3), lds128(sm, o0)};
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const unsigned char* b0 = sm + ((it * 2) & 3) * 16384u;
            const unsigned char* b1 = sm + ((it * 2 + 1) & 3) * 16384u;
            if (V != V_SEL_IADD3_NOLDS) {
                r0 = BlockRows{lds128(b0, o0), lds128(b0, o1), lds128(b0, o2), lds128(b0, o3)};
                r1 = BlockRows{lds128(b1, o0), lds128(b1, o1), lds128(b1, o2), lds128(b1, o3)};
            } else { r0.pu.x ^= x.gt[0][0]; r1.pv.y ^= x.lt[3][2]; }    // loop-carried, so nothing is hoisted
            if (V == V_PRMT_IADD3) step_prmt(x, r0, r1); else step_gt_lt(x, r0, r1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc ^= x.gt[j][p] ^ x.lt[j][p];
    } else if (V == V_TP_IADD3 || V == V_TP_IADD3_NOLDS) {
        TPCounters c;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int p = 0; p < 8; ++p) { c.gt[j][p] = 0; c.lt[j][p] = 0; }
        uint4 pu0[2], qu0[2], pu1[2], qu1[2], pv0, qv0, pv1, qv1;
        pu0[0] = lds128(sm, o0); pu0[1] = lds128(sm, o0 + 16); qu0[0] = lds128(sm, o1); qu0[1] = lds128(sm, o1 + 16); pv0 = lds128(sm, o2); qv0 = lds128(sm, o3);
        pu1[0] = lds128(sm, o1); pu1[1] = lds128(sm, o2 + 16); qu1[0] = lds128(sm, o3); qu1[1] = lds128(sm, o0 + 32); pv1 = lds128(sm, o2 + 32); qv1 = lds128(sm, o3 + 32);
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const unsigned char* b0 = sm + ((it * 2) & 3) * 16384u;
            const unsigned char* b1 = sm + ((it * 2 + 1) & 3) * 16384u;
            if (V == V_TP_IADD3) {
                pu0[0] = lds128(b0, o0); pu0[1] = lds128(b0, o0 + 16); qu0[0] = lds128(b0, o1); qu0[1] = lds128(b0, o1 + 16); pv0 = lds128(b0, o2); qv0 = lds128(b0, o3);
                pu1[0] = lds128(b1, o0); pu1[1] = lds128(b1, o0 + 16); qu1[0] = lds128(b1, o1); qu1[1] = lds128(b1, o1 + 16); pv1 = lds128(b1, o2); qv1 = lds128(b1, o3);
            } else { pu0[0].x ^= c.gt[0][0]; pv1.y ^= c.lt[3][2]; }
            step_tp(c, pu0, qu0, pv0, qv0, pu1, qu1, pv1, qv1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc ^= c.gt[j][p] ^ c.lt[j][p];
    } else {
        HCounters x;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int p = 0; p < 4; ++p) { x.gt[j][p] = __float2half2_rn(0.f); x.lt[j][p] = __float2half2_rn(0.f); }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const unsigned char* b0 = sm + ((it * 2) & 3) * 16384u;
            const unsigned char* b1 = sm + ((it * 2 + 1) & 3) * 16384u;
            BlockRows r0{lds128(b0, o0), lds128(b0, o1), lds128(b0, o2), lds128(b0, o3)};
            BlockRows r1{lds128(b1, o0), lds128(b1, o1), lds128(b1, o2), lds128(b1, o3)};
            step_hadd(x, r0, r1);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc ^= *reinterpret_cast<uint32_t*>(&x.gt[j][p]) ^ *reinterpret_cast<uint32_t*>(&x.lt[j][p]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void run(int nsm, int threads, double clk_hz, uint32_t* out, uint32_t* in) {
    CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k<V><<<nsm, threads, 65536>>>(out, in);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<V><<<nsm, threads, 65536>>>(out, in); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double hset2 = double(threads / 32) * ITERS * 128.0;     // HSET2 warp-instr per SM
    printf("%-75s thr=%4d time=%8.3f ms  HSET2/clk/SM=%6.3f (pipe peak 2.0)\n", vname[V], threads, best, hset2 / (best * 1e-3 * clk_hz));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max SM clock %d kHz\n", p.name, p.multiProcessorCount, clk_khz);
    uint32_t *out, *in; CK(cudaMalloc(&out, 4096 * 1024 * 4)); CK(cudaMalloc(&in, 1024));
    uint32_t h[256]; for (int i = 0; i < 256; i++) h[i] = 0x40004000u + ((i * 37) & 15) * 0x04000400u + ((i * 11) & 7) * 0x0400u;
    CK(cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice));
    for (int threads : {256, 384, 512}) {
        run<V_SEL_IADD3>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
        run<V_TP_IADD3>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
        run<V_SEL_HADD2>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
        run<V_SEL_IADD3_NOLDS>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
        run<V_TP_IADD3_NOLDS>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
        run<V_PRMT_IADD3>(p.multiProcessorCount, threads, clk_khz * 1e3, out, in);
    }
    return 0;
}
