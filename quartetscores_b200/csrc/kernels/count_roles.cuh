// count_roles.cuh — the per-(thread, tree) inner steps of the counting kernels.
//
// Four-point test in "fixed pair" form (derivation in count_rows.cuh).  For a fixed taxon pair (p,q) and
// G_pq(t) = D[q][t] - D[p][t]:   G(u) > G(v)  <=>  the tree displays up|vq.
//
// Instruction mix, chosen from measurements on B200 (profiles/r01_b_ubench_pipes.txt,
// profiles/r01_c_ubench_loop.txt and the ncu captures next to them):
//   * compare: HSET2.BF.GT / .LT on packed fp16x2 (distances are small integers, exact in fp16): two
//     quartets per lane, result 1.0 / 0.0 per half.  A missing taxon is NaN -> ordered compare false.
//     HSET2 runs on the ALU pipe (0.5 warp-instr/clk/SMSP).
//   * accumulate: HADD2 on the fp16 pipe (0.5 warp-instr/clk/SMSP).  Counters start at -2048 so that a
//     chunk may hold 4096 trees (fp16 represents every integer in [-2048, 2048] exactly).
//   The alternative "integer mask + one three-input IADD3 per two masks" needs fewer instructions and looked
//   better in an isolated microbenchmark, but IADD3 shares the ALU pipe with HSET2: in the real loop ncu
//   shows pipe_alu at 97 % and 1.22 compares/clk/SM, against 1.46 for HSET2.BF + HADD2 (ALU 79 %,
//   fp16 83 %, issue 84 %) — so the two half-rate pipes are loaded evenly on purpose.
#pragma once
#include "common.cuh"

namespace qs {

constexpr int QS_MAX_CHUNK_TREES = 4096;

#ifndef QS_INT_COUNTERS
#define QS_INT_COUNTERS 0
#endif

#if QS_INT_COUNTERS
// Tuning build: integer-mask compare (HSET2 -> 0xFFFF per true half) and a TWO-input integer subtract per compare, which
// ptxas places on the FMA pipe as IMAD.IADD: 1.71 compares+accumulates/clk/SM in isolation against 1.55 for
// HSET2.BF + HADD2 (tools/ubench_mix2.cu, profiles/r01_w_ubench_mix2.txt).  Subtracting 0xFFFF from a 16-bit half adds 1
// to it and borrows 1 from the upper half, so after L low-half and H high-half hits acc = L + 65536 (H - L) mod 2^32.
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
