// count_roles.cuh — the per-(thread, tree pair) inner steps shared by both counting kernels.
//
// Four-point test in "fixed pair" form (derivation in count_items.cuh).  For a fixed taxon pair (p,q)
// and G_pq(t) = D[q][t] - D[p][t]:   G(u) > G(v)  <=>  the tree displays up|vq.
//
// Instruction mix (measured on B200, profiles/r01_b_ubench_pipes.txt):
//   * compare: HSET2.GT / HSET2.LT on packed fp16x2 (distances are small integers, exact in fp16) with
//     an INTEGER MASK result (0xFFFF per true half).  A missing taxon is NaN -> ordered compare false.
//     HSET2 issues at 2.0 warp-instr/clk/SM; it is the pipe this kernel is bound by.
//   * accumulate: ONE three-input integer add per TWO masks (acc = acc - m0 - m1, SASS IADD3 with both
//     operands negated).  Subtracting 0xFFFF from a 16-bit half adds 1 to it and borrows from the
//     upper half, so after L low-half hits and H high-half hits   acc = L + 65536*(H - L)  (mod 2^32),
//     which decode() inverts exactly while L,H <= 65535.  The two masks of one add belong to the same
//     counter and two different gene trees, so trees are processed in pairs.
//     HSET2 x2 + IADD3 sustains 2.96 warp-instr/clk/SM = 1.97 compares/clk/SM, against 1.6 for the
//     fp16 HSET2.BF + HADD2 pair used before (issue-limited at 3.2) — and the counters now hold 65535
//     trees instead of 2048.
#pragma once
#include "common.cuh"

namespace qs {

constexpr int QS_MAX_CHUNK_TREES = 65534;   // even, <= 65535: capacity of one 16-bit counter half

struct XCounters { uint32_t gt[8][4], lt[8][4]; };   // 8 (b) x 8 (a, packed in pairs): G(a)>G(b), G(a)<G(b)
struct GCounters { uint32_t gt[8][4]; };             // 8 x 8, G(u)>G(v) only

__device__ __forceinline__ void zero(XCounters& x) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) { x.gt[j][p] = 0u; x.lt[j][p] = 0u; }
}
__device__ __forceinline__ void zero(GCounters& y) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) y.gt[j][p] = 0u;
}

// packed counter -> (hits in the low half, hits in the high half)
__device__ __forceinline__ void decode(uint32_t acc, uint32_t& lo, uint32_t& hi) {
    lo = acc & 0xFFFFu;
    hi = ((acc >> 16) + lo) & 0xFFFFu;
}

__device__ __forceinline__ void sub4(__half2 (&g)[4], const uint4& hi, const uint4& lo) {
    g[0] = __hsub2(as_h2(hi.x), as_h2(lo.x)); g[1] = __hsub2(as_h2(hi.y), as_h2(lo.y));
    g[2] = __hsub2(as_h2(hi.z), as_h2(lo.z)); g[3] = __hsub2(as_h2(hi.w), as_h2(lo.w));
}

// One tree's operands of an 8 x 8 block: rows p,q of D at the 8 "u" columns and at the 8 "v" columns.
struct BlockRows { uint4 pu, qu, pv, qv; };

// role X (pair (c,d) fixed, u = a, v = b): G(a) > G(b) -> ac|bd (slot 1),  G(a) < G(b) -> ad|bc (slot 2)
__device__ __forceinline__ void step_gt_lt(XCounters& x, const BlockRows& r0, const BlockRows& r1) {
    __half2 u0[4], v0[4], u1[4], v1[4];
    sub4(u0, r0.qu, r0.pu); sub4(v0, r0.qv, r0.pv);
    sub4(u1, r1.qu, r1.pu); sub4(v1, r1.qv, r1.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b0 = (j & 1) ? __high2half2(v0[j >> 1]) : __low2half2(v0[j >> 1]);
        const __half2 b1 = (j & 1) ? __high2half2(v1[j >> 1]) : __low2half2(v1[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            x.gt[j][p] = x.gt[j][p] - __hgt2_mask(u0[p], b0) - __hgt2_mask(u1[p], b1);
            x.lt[j][p] = x.lt[j][p] - __hlt2_mask(u0[p], b0) - __hlt2_mask(u1[p], b1);
        }
    }
}

// G(u) > G(v) only: role Y (pair (b,c) fixed, u = a, v = d: ab|cd, slot 0) and the diagonal blocks of role X
__device__ __forceinline__ void step_gt(GCounters& y, const BlockRows& r0, const BlockRows& r1) {
    __half2 u0[4], v0[4], u1[4], v1[4];
    sub4(u0, r0.qu, r0.pu); sub4(v0, r0.qv, r0.pv);
    sub4(u1, r1.qu, r1.pu); sub4(v1, r1.qv, r1.pv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b0 = (j & 1) ? __high2half2(v0[j >> 1]) : __low2half2(v0[j >> 1]);
        const __half2 b1 = (j & 1) ? __high2half2(v1[j >> 1]) : __low2half2(v1[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) y.gt[j][p] = y.gt[j][p] - __hgt2_mask(u0[p], b0) - __hgt2_mask(u1[p], b1);
    }
}

// diagonal block of role X (u and v are the same 8 taxa): G(x) > G(y) for all ordered pairs; one load per row
__device__ __forceinline__ void step_gt_diag(GCounters& y, const uint4& p0, const uint4& q0, const uint4& p1, const uint4& q1) {
    __half2 u0[4], u1[4];
    sub4(u0, q0, p0);
    sub4(u1, q1, p1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 b0 = (j & 1) ? __high2half2(u0[j >> 1]) : __low2half2(u0[j >> 1]);
        const __half2 b1 = (j & 1) ? __high2half2(u1[j >> 1]) : __low2half2(u1[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) y.gt[j][p] = y.gt[j][p] - __hgt2_mask(u0[p], b0) - __hgt2_mask(u1[p], b1);
    }
}

}  // namespace qs
