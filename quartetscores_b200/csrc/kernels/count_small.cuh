// count_small.cuh — quartet counting kernel for reference trees whose whole n x n distance matrix
// fits in shared memory several times over (n <= ~230): BASELINE config 1/2 shapes.
//
// Replaces the reference hot loop QuartetCounterLookup::updateQuartetsThreeClades
// (src/QuartetCounterLookup.hpp:66-106) + countQuartets (:197-238).  Instead of enumerating clades per
// tree and scattering increments into an n^4 table, every thread OWNS a fixed set of quartets, keeps
// their topology counters in registers across all gene trees, and decides each (quartet, tree) with the
// four-point condition on the tree's distance matrix D staged in shared memory by TMA bulk copies.
//
// Four-point test in "fixed pair" form.  For taxa p,q define G_pq(t) = D[q][t] - D[p][t].  For a tree
// metric, G_pq(u) > G_pq(v)  <=>  D_up + D_vq < D_uq + D_vp  <=>  the tree displays up|vq (the two
// larger pair sums of a tree metric are equal, so one strict inequality decides the topology;
// ties = unresolved or the third topology).  With sorted ids a<b<c<d and the table slots
// 0 = ab|cd, 1 = ac|bd, 2 = ad|bc (src/quartet_lookup_table.hpp:87-111):
//   role X, pair (c,d) fixed:  G_cd(a) > G_cd(b) -> slot 1,   G_cd(a) < G_cd(b) -> slot 2
//   role Y, pair (b,c) fixed:  G_bc(a) > G_bc(d) -> slot 0
// A missing taxon makes D = NaN in its row/column, G = NaN, and every ordered compare false, so such
// (quartet, tree) pairs count nothing — exactly the reference, where absent taxa are never enumerated.
//
// Arithmetic: distances are small integers, exact in fp16 (<= 2048, checked by the caller), so two
// evaluations are packed per 32-bit lane: HSET2.BF (compare -> 1.0/0.0, ALU pipe) + HADD2 (accumulate,
// FMA pipe).  Measured on B200 (profiles/r01_ubench_pipes.txt) that pair of instructions issues at
// 4 warp-instr/clk/SM = the SM's full issue rate.  fp16 counters are exact up to 2048, so trees are
// processed in chunks of <= 2048 and flushed to a uint32 workspace with red.global.add.
//
// Work decomposition: an X item is (c,d) x 8 consecutive a x 8 consecutive b; a Y item is (b,c) x 8 a
// x 8 d.  Each thread owns one X item and one Y item (64 quartets each, 64+32 packed counter
// registers).  Items are enumerated over the whole (sharded) quartet space — see host-side tables
// PX/PY/CD — so no thread is idle; only 8-aligned chunk boundaries waste lanes.
#pragma once
#include "common.cuh"
#include "count_roles.cuh"

namespace qs {

struct CountSmallArgs {
    const __half* D;        // [m][n][n_pad] fp16, NaN = missing
    uint32_t* ws;           // [(rank - rank_base)*3 + slot] uint32 workspace (zeroed by caller)
    const int32_t* PX;      // [n+1] #X items with fixed c' < c
    const int32_t* PY;      // [n+1] #Y items with fixed b' < b
    const int32_t* CD;      // [n+1] #d-chunks of Y over c' < c
    uint64_t rank_base;
    int n, n_pad, m;
    int d_begin, d_end;     // shard: quartets with d in [d_begin, d_end)
    int NX, NY;
    int n_item_blocks, n_tree_chunks, chunk_trees;
    int trees_per_stage;
    uint32_t tree_bytes;    // n * n_pad * 2
};

constexpr int CS_STAGES = 2;

__device__ __forceinline__ int cs_upper_bound(const int32_t* __restrict__ P, int lo, int hi, int key) {
    // largest x in [lo,hi) with P[x] <= key  (P non-decreasing)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (P[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) qs_count_small_kernel(const CountSmallArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* bufs = smem + 128;
    const int tid = threadIdx.x;
    const uint32_t stage_bytes = a.trees_per_stage * a.tree_bytes;

    if (tid == 0) {
        for (int s = 0; s < CS_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;   // bit s = parity to wait for on full[s]

    const int n_tasks = a.n_item_blocks * a.n_tree_chunks;
    for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
        const int ib_idx = task % a.n_item_blocks;
        const int tc = task / a.n_item_blocks;
        const int t0 = tc * a.chunk_trees;
        const int t1 = min(a.m, t0 + a.chunk_trees);

        // ---- decode this thread's items ------------------------------------------------------
        // X item: (c, d, ia, ib)
        int xc = 0, xd = 0, xia = 0, xib = 0; bool xvalid = false;
        {
            int i = ib_idx * THREADS + tid;
            if (i < a.NX) {
                xvalid = true;
                xc = cs_upper_bound(a.PX, 0, a.n, i);
                int r = i - a.PX[xc];
                int nb = (xc + 7) >> 3;
                int K = nb * (nb + 1) / 2;
                int dlo = max(xc + 1, a.d_begin);
                xd = dlo + r / K;
                int rr = r % K;
                xib = (int)((sqrtf(8.f * rr + 1.f) - 1.f) * 0.5f);
                while ((xib + 1) * (xib + 2) / 2 <= rr) ++xib;
                while (xib * (xib + 1) / 2 > rr) --xib;
                xia = rr - xib * (xib + 1) / 2;
            }
        }
        // Y item: (b, c, ia, id)
        int yb = 0, yc = 0, yia = 0, yid = 0; bool yvalid = false;
        {
            int i = ib_idx * THREADS + tid;
            if (i < a.NY) {
                yvalid = true;
                yb = cs_upper_bound(a.PY, 0, a.n, i);
                int r = i - a.PY[yb];
                int na = (yb + 7) >> 3;
                int q = r / na;
                yia = r % na;
                int target = a.CD[yb + 1] + q;
                yc = cs_upper_bound(a.CD, yb + 1, a.n, target);
                int dlo = max(yc + 1, a.d_begin);
                yid = (dlo >> 3) + (target - a.CD[yc]);
            }
        }
        const uint32_t rowb = (uint32_t)a.n_pad * 2u;
        const uint32_t oXca = xc * rowb + xia * 16u, oXda = xd * rowb + xia * 16u;
        const uint32_t oXcb = xc * rowb + xib * 16u, oXdb = xd * rowb + xib * 16u;
        const uint32_t oYba = yb * rowb + yia * 16u, oYca = yc * rowb + yia * 16u;
        const uint32_t oYbd = yb * rowb + yid * 16u, oYcd = yc * rowb + yid * 16u;

        XCounters xc_; YCounters yc_;
        zero(xc_); zero(yc_);

        // ---- stream the trees of this chunk through shared memory ------------------------------
        const int ntrees = t1 - t0;
        const int nst = (ntrees + a.trees_per_stage - 1) / a.trees_per_stage;
        __syncthreads();   // previous task's reads of the buffers are done
        if (tid == 0) {
            for (int s = 0; s < CS_STAGES && s < nst; ++s) {
                int nt = min(a.trees_per_stage, ntrees - s * a.trees_per_stage);
                mbar_expect_tx(&full[s], nt * a.tree_bytes);
                bulk_g2s(bufs + s * stage_bytes, a.D + (size_t)(t0 + s * a.trees_per_stage) * (a.tree_bytes / 2), nt * a.tree_bytes, &full[s]);
            }
        }
        for (int s = 0; s < nst; ++s) {
            const int buf = s % CS_STAGES;
            mbar_wait(&full[buf], (phase >> buf) & 1u);
            phase ^= (1u << buf);
            const int nt = min(a.trees_per_stage, ntrees - s * a.trees_per_stage);
            const unsigned char* base = bufs + buf * stage_bytes;
            for (int tt = 0; tt < nt; ++tt, base += a.tree_bytes) {
                role_x_step(xc_, lds128(base, oXca), lds128(base, oXda), lds128(base, oXcb), lds128(base, oXdb));
                role_y_step(yc_, lds128(base, oYba), lds128(base, oYca), lds128(base, oYbd), lds128(base, oYcd));
            }
            __syncthreads();   // everyone is done reading this buffer
            if (tid == 0 && s + CS_STAGES < nst) {
                const int s2 = s + CS_STAGES;
                int nt2 = min(a.trees_per_stage, ntrees - s2 * a.trees_per_stage);
                mbar_expect_tx(&full[buf], nt2 * a.tree_bytes);
                bulk_g2s(bufs + buf * stage_bytes, a.D + (size_t)(t0 + s2 * a.trees_per_stage) * (a.tree_bytes / 2), nt2 * a.tree_bytes, &full[buf]);
            }
        }

        // ---- flush the fp16 counters (exact, <= 2048) into the uint32 workspace -------------------
        if (xvalid) {
            const uint64_t rcd = binom4((uint64_t)xd) + binom3((uint64_t)xc) - a.rank_base;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = xib * 8 + j;
                if (b >= xc) continue;
                const uint64_t rb = rcd + (uint64_t)b * (b - 1) / 2;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 v1 = __half22float2(xc_.s1[j][p]), v2 = __half22float2(xc_.s2[j][p]);
                    const int a0 = xia * 8 + 2 * p;
                    if (a0 < b) {
                        uint32_t* w = a.ws + (rb + a0) * 3;
                        if (v1.x != 0.f) atomicAdd(w + 1, (uint32_t)v1.x);
                        if (v2.x != 0.f) atomicAdd(w + 2, (uint32_t)v2.x);
                    }
                    if (a0 + 1 < b) {
                        uint32_t* w = a.ws + (rb + a0 + 1) * 3;
                        if (v1.y != 0.f) atomicAdd(w + 1, (uint32_t)v1.y);
                        if (v2.y != 0.f) atomicAdd(w + 2, (uint32_t)v2.y);
                    }
                }
            }
        }
        if (yvalid) {
            const int dlo = max(yc + 1, a.d_begin);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d = yid * 8 + j;
                if (d < dlo || d >= a.d_end) continue;
                const uint64_t rb = binom4((uint64_t)d) + binom3((uint64_t)yc) + (uint64_t)yb * (yb - 1) / 2 - a.rank_base;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 v0 = __half22float2(yc_.s0[j][p]);
                    const int a0 = yia * 8 + 2 * p;
                    if (a0 < yb && v0.x != 0.f) atomicAdd(a.ws + (rb + a0) * 3, (uint32_t)v0.x);
                    if (a0 + 1 < yb && v0.y != 0.f) atomicAdd(a.ws + (rb + a0 + 1) * 3, (uint32_t)v0.y);
                }
            }
        }
    }
}

// uint32 workspace -> CINT table (QuartetLookupTable layout), elementwise
template <typename CINT>
__global__ void qs_narrow_kernel(const uint32_t* __restrict__ ws, CINT* __restrict__ table, uint64_t n_elems) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_elems; i += stride) table[i] = (CINT)ws[i];
}

}  // namespace qs
