// ubench.cuh — live measurement of the issue rates that bound the counting kernel (roofline denominators).
// compare mix: 2 x HSET2 (fp16x2 compare -> integer mask) + 1 x IADD3 (three-input add of both masks), the counting
// kernel's inner-loop triple; the reported rate counts the HSET2 lane-ops only (the pipe the kernel is bound by).
// int32 mix: LOP3 on the ALU pipe (the classic INT32 lane rate; IADD3 issues at the same rate).
#pragma once
#include "common.cuh"

namespace qs {

template <int MIX>
__global__ void __launch_bounds__(512) qs_ubench_kernel(uint32_t* out, const uint32_t* in, int iters) {
    constexpr int NCH = 16;
    uint32_t acc[NCH], x[NCH];
    const uint32_t y = in[threadIdx.x & 31];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { acc[i] = in[(threadIdx.x + i) & 63]; x[i] = in[(threadIdx.x * 3 + i) & 63]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MIX == 0) {
                // the compares read other chains' accumulators, so ptxas cannot hoist them out of the loop
                if (i & 1) {
                    uint32_t m0, m1;
                    asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m0) : "r"(x[i]), "r"(acc[(i + 5) % NCH]));
                    asm volatile("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m1) : "r"(x[i - 1]), "r"(acc[(i + 6) % NCH]));
                    asm volatile("{.reg .s32 t; add.s32 t, %1, %2; sub.s32 %0, %0, t;}" : "+r"(acc[i]) : "r"(m0), "r"(m1));
                }
            } else {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(acc[(i + 5) % NCH]), "r"(y));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

inline cudaError_t ubench_alu_peak(int num_sms, cudaStream_t stream, double* hset2_laneops, double* int_laneops) {
    uint32_t *out = nullptr, *in = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&out, (size_t)num_sms * 512 * 4)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&in, 256)) != cudaSuccess) { cudaFree(out); return e; }
    uint32_t h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0x3c003c00u + i * 0x00010001u;
    cudaMemcpyAsync(in, h, 256, cudaMemcpyHostToDevice, stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192;
    double res[2] = {0, 0};
    for (int mix = 0; mix < 2; ++mix) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {     // rep 0 is the warm-up
            cudaEventRecord(e0, stream);
            if (mix == 0) qs_ubench_kernel<0><<<num_sms, 512, 0, stream>>>(out, in, iters);
            else qs_ubench_kernel<1><<<num_sms, 512, 0, stream>>>(out, in, iters);
            cudaEventRecord(e1, stream);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        if (e != cudaSuccess) break;
        const double instr_per_thread = (double)iters * 16;   // mix 0: 16 HSET2 (+ 8 IADD3, not counted) per iteration
        res[mix] = instr_per_thread * 512.0 * num_sms / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(in);
    if (e == cudaSuccess) e = cudaGetLastError();
    *hset2_laneops = res[0]; *int_laneops = res[1];
    return e;
}

}  // namespace qs
