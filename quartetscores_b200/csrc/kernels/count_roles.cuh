// count_roles.cuh — the per-(thread, tree) inner step shared by both counting kernels.
// See count_small.cuh for the derivation (fixed-pair form of the four-point condition).
#pragma once
#include "common.cuh"

namespace qs {

struct XCounters { __half2 s1[8][4], s2[8][4]; };   // slot 1 / slot 2 of 8 (b) x 8 (a, packed in pairs)
struct YCounters { __half2 s0[8][4]; };             // slot 0 of 8 (d) x 8 (a)

__device__ __forceinline__ void zero(XCounters& x) {
    const __half2 z = __float2half2_rn(0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) { x.s1[j][p] = z; x.s2[j][p] = z; }
}
__device__ __forceinline__ void zero(YCounters& y) {
    const __half2 z = __float2half2_rn(0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) y.s0[j][p] = z;
}

__device__ __forceinline__ void sub4(__half2 (&g)[4], const uint4& hi, const uint4& lo) {
    g[0] = __hsub2(as_h2(hi.x), as_h2(lo.x)); g[1] = __hsub2(as_h2(hi.y), as_h2(lo.y));
    g[2] = __hsub2(as_h2(hi.z), as_h2(lo.z)); g[3] = __hsub2(as_h2(hi.w), as_h2(lo.w));
}

// role X, pair (c,d) fixed: G_cd(t) = D[d][t] - D[c][t];  G(a) > G(b) -> ac|bd (slot 1),  G(a) < G(b) -> ad|bc (slot 2)
// ca/da = rows c,d at the 8 a-columns, cb/db = rows c,d at the 8 b-columns
__device__ __forceinline__ void role_x_step(XCounters& x, const uint4& ca, const uint4& da, const uint4& cb, const uint4& db) {
    __half2 ga[4], gb[4];
    sub4(ga, da, ca);
    sub4(gb, db, cb);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 bj = (j & 1) ? __high2half2(gb[j >> 1]) : __low2half2(gb[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            x.s1[j][p] = __hadd2(x.s1[j][p], __hgt2(ga[p], bj));
            x.s2[j][p] = __hadd2(x.s2[j][p], __hlt2(ga[p], bj));
        }
    }
}

// role Y, pair (b,c) fixed: G_bc(t) = D[c][t] - D[b][t];  G(a) > G(d) -> ab|cd (slot 0)
__device__ __forceinline__ void role_y_step(YCounters& y, const uint4& ba, const uint4& ca, const uint4& bd, const uint4& cd) {
    __half2 ga[4], gd[4];
    sub4(ga, ca, ba);
    sub4(gd, cd, bd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const __half2 dj = (j & 1) ? __high2half2(gd[j >> 1]) : __low2half2(gd[j >> 1]);
#pragma unroll
        for (int p = 0; p < 4; ++p) y.s0[j][p] = __hadd2(y.s0[j][p], __hgt2(ga[p], dj));
    }
}

}  // namespace qs
