// count_roles.cuh — the per-(thread, tree) inner steps of the counting kernels.
//
// Four-point test in "fixed pair" form (derivation in count_rows.cuh).  For a fixed taxon pair (p,q) and
// G_pq(t) = D[q][t] - D[p][t]:   G(u) > G(v)  <=>  the tree displays up|vq.
//
// Instruction mix, chosen from measurements on B200 (tools/ubench_pipes.cu, tools/ubench_mix2.cu and the tuning builds timed
// in profiles/r01_x_*, r01_zm_*):
//   * compare: HSET2.GT / .LT on packed fp16x2 (distances are small integers, exact in fp16): two quartets per lane, result
//     an integer mask, 0xFFFF per true half.  A missing taxon is NaN -> ordered compare false.  ALU pipe, half rate.
//   * accumulate: one TWO-input integer subtract per compare, acc -= mask, which ptxas places on the FMA pipe as
//     IMAD.IADD.  HSET2 + IMAD.IADD issue at 3.42 of 4 warp-instr/clk/SM in isolation (1.71 compares+accumulates), against
//     3.10 (1.55) for the fp16 pair HSET2.BF + HADD2 that rounds 1a-1y used, and in the kernel it is 4-6 % faster — provided
//     the tree loop is NOT unrolled: with two trees per iteration ptxas fuses the two subtracts of a counter into one
//     three-input IADD3, which runs on the ALU pipe next to the HSET2s and loses 10 % (the "integer counters are slower"
//     result of rounds 1c and 1x was that fusion).  QS_INT_COUNTERS=0 selects the fp16 pair again.
//   Subtracting 0xFFFF from a 16-bit half adds 1 to it and borrows 1 from the upper half, so after L low-half and H
//   high-half hits acc = L + 65536 (H - L) mod 2^32: exact while a chunk holds <= 65535 trees (QS_MAX_CHUNK_TREES = 4096).
#pragma once
#include "common.cuh"

namespace qs {

constexpr int QS_MAX_CHUNK_TREES = 4096;

#ifndef QS_INT_COUNTERS
#define QS_INT_COUNTERS 1
#endif

#if QS_INT_COUNTERS
// integer-mask compare + two-input subtract (see the header comment)
typedef uint32_t ctr_t;
__device__ __forceinline__ ctr_t ctr_zero() { return 0u; }
#if QS_INT_COUNTERS == 2      // keep every subtract a separate two-input instruction (ptxas otherwise fuses two trees into one IADD3)
__device__ __forceinline__ void acc_sub(ctr_t& acc, uint32_t m) { asm volatile("sub.s32 %0, %0, %1;" : "+r"(acc) : "r"(m)); }
#else
__device__ __forceinline__ void acc_sub(ctr_t& acc, uint32_t m) { acc -= m; }
#endif
__device__ __forceinline__ void acc_gt(ctr_t& acc, __half2 u, __half2 b) { acc_sub(acc, __hgt2_mask(u, b)); }
__device__ __forceinline__ void acc_lt(ctr_t& acc, __half2 u, __half2 b) { acc_sub(acc, __hlt2_mask(u, b)); }
__device__ __forceinline__ void decode(ctr_t acc, uint32_t& lo, uint32_t& hi) {
    lo = acc & 0xffffu;
    hi = ((acc >> 16) + lo) & 0xffffu;
}
#else
constexpr float QS_COUNTER_BIAS = 2048.f;
typedef __half2 ctr_t;
__device__ __forceinline__ ctr_t ctr_zero() { return __float2half2_rn(-QS_COUNTER_BIAS); }
__device__ __forceinline__ void acc_gt(ctr_t& acc, __half2 u, __half2 b) { acc = __hadd2(acc, __hgt2(u, b)); }
__device__ __forceinline__ void acc_lt(ctr_t& acc, __half2 u, __half2 b) { acc = __hadd2(acc, __hlt2(u, b)); }
// packed counter -> (hits of the low half, hits of the high half)
__device__ __forceinline__ void decode(ctr_t acc, uint32_t& lo, uint32_t& hi) {
    const float2 f = __half22float2(acc);
    lo = (uint32_t)(f.x + QS_COUNTER_BIAS);
    hi = (uint32_t)(f.y + QS_COUNTER_BIAS);
}
#endif

struct XCounters { ctr_t gt[8][4], lt[8][4]; };   // 8 (v) x 8 (u, packed in pairs): G(u)>G(v), G(u)<G(v)
struct GCounters { ctr_t gt[8][4]; };             // 8 x 8, G(u)>G(v) only

__device__ __forceinline__ void zero(XCounters& x) {
    const ctr_t z = ctr_zero();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) { x.gt[j][p] = z; x.lt[j][p] = z; }
}
__device__ __forceinline__ void zero(GCounters& y) {
    const ctr_t z = ctr_zero();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) y.gt[j][p] = z;
}

__device__ __forceinline__ void sub4(__half2 (&g)[4], const uint4& hi, const uint4& lo) {
    g[0] = __hsub2(as_h2(hi.x), as_h2(lo.x)); g[1] = __hsub2(as_h2(hi.y), as_h2(lo.y));
    g[2] = __hsub2(as_h2(hi.z), as_h2(lo.z)); g[3] = __hsub2(as_h2(hi.w), as_h2(lo.w));
}

// One tree's operands of an 8 x 8 block: rows p,q of D at the 8 "u" columns and at the 8 "v" columns.
struct BlockRows { uint4 pu, qu, pv, qv; };

// role X (pair (c,d) fixed, u = a, v = b): G(a) > G(b) -> ac|bd (slot 1),  G(a) < G(b) -> ad|bc (slot 2)
__device__ __forceinline__ void step_gt_lt(XCounters& x, const BlockRows& r) {
    __half2 u[4], v[4];
    sub4(u, r.qu, r.pu);
    sub4(v, r.qv, r.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b = (j & 1) ? __high2half2(v[j >> 1]) : __low2half2(v[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc_gt(x.gt[j][p], u[p], b);
            acc_lt(x.lt[j][p], u[p], b);
        }
    }
}

// the same with G already formed, over the first JM of the 8 "v" taxa (JM < 8: the ragged b-block of role X, count_rows.cuh)
struct BlockG { __half2 u[4], v[4]; };
template <int JM>
__device__ __forceinline__ void step_gt_lt_g(XCounters& x, const BlockG& g) {
#pragma unroll
    for (int j = 0; j < JM; ++j) {
        const __half2 b = (j & 1) ? __high2half2(g.v[j >> 1]) : __low2half2(g.v[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            acc_gt(x.gt[j][p], g.u[p], b);
            acc_lt(x.lt[j][p], g.u[p], b);
        }
    }
}

// G(u) > G(v) only: role Y (pair (b,c) fixed, u = a, v = d: ab|cd, slot 0)
__device__ __forceinline__ void step_gt(GCounters& y, const BlockRows& r) {
    __half2 u[4], v[4];
    sub4(u, r.qu, r.pu);
    sub4(v, r.qv, r.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b = (j & 1) ? __high2half2(v[j >> 1]) : __low2half2(v[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) acc_gt(y.gt[j][p], u[p], b);
    }
}

// diagonal block of role X (u and v are the same 8 taxa): G(x) > G(y) for all ordered pairs; one load per row
__device__ __forceinline__ void step_gt_diag(GCounters& y, const uint4& p, const uint4& q) {
    __half2 u[4];
    sub4(u, q, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b = (j & 1) ? __high2half2(u[j >> 1]) : __low2half2(u[j >> 1]);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) acc_gt(y.gt[j][pp], u[pp], b);
    }
}

}  // namespace qs
