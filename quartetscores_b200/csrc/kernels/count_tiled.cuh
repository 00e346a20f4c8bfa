// count_tiled.cuh — quartet counting kernel for any n: Cartesian 4-D tiles of the quartet space with the
// needed sub-blocks of every gene tree's distance matrix staged in shared memory by TMA tensor loads.
//
// Same arithmetic as count_small.cuh (fixed-pair four-point test, fp16x2 HSET2 + HADD2, counters in
// registers across all trees).  A tile is A x B x C x D = 16 x 16 x 16 x 8 taxon ids (32,768 quartets);
// its 512 threads each own one X item ((c,d) fixed, 8 a x 8 b) and one Y item ((b,c) fixed, 8 a x 8 d).
// Per tree the tile needs seven sub-blocks of D (rows x cols): CxA, DxA, CxB, DxB (role X) and
// BxA, BxD, CxD (+CxA again) (role Y) = 1,280 fp16 = 2,560 bytes for 32,768 evaluations (0.08 B/eval),
// fetched with cp.async.bulk.tensor.3d (UTMALDG) from the [tree][row][col] tensor.
//
// After the last tree the counters go through a per-CTA scratch (global, L2-resident) so that the
// epilogue can (a) write whole table entries, 16 consecutive a = 96 contiguous bytes per half-warp, in
// QuartetLookupTable layout (src/quartet_lookup_table.hpp:135-212) and/or (b) score the quartets on the
// spot (table-free mode, the -s analogue) with the same pair aggregation as kernels/score.cuh.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "count_roles.cuh"
#include "score.cuh"

namespace qs {

constexpr int CT_THREADS = 512;
constexpr int CT_TA = 16, CT_TB = 16, CT_TC = 16, CT_TD = 8;
constexpr int CT_TILE_Q = CT_TA * CT_TB * CT_TC * CT_TD;        // 32768 quartets
constexpr int CT_TREE_BYTES = 2560;
constexpr int CT_TPS = 4;                                        // trees per pipeline stage
constexpr int CT_STAGES = 4;
// block offsets inside one tree's 2560-byte record
constexpr int CT_OFF_CA = 0, CT_OFF_DA = 512, CT_OFF_CB = 768, CT_OFF_DB = 1280, CT_OFF_BA = 1536, CT_OFF_BD = 2048, CT_OFF_CD = 2304;

struct CountTiledArgs {
    const ushort4* tiles;       // (iA, iB, iC, jD): A/B/C ranges start at 16*i, D range at 8*j
    int n_tiles;
    int n, m;
    int d_begin, d_end;
    uint64_t rank_base;
    uint32_t* scratch;          // [gridDim.x][CT_TILE_Q*3]
    void* table;                // CINT table or nullptr
    int cint_bytes;
    int fused_score;            // table-free: aggregate into sa.pair_*
    ScoreArgs sa;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

template <typename CINT>
__device__ __forceinline__ void ct_write_entry(void* table, uint64_t idx, uint32_t v0, uint32_t v1, uint32_t v2) {
    CINT* t = reinterpret_cast<CINT*>(table) + idx * 3;
    t[0] = (CINT)v0; t[1] = (CINT)v1; t[2] = (CINT)v2;
}

__global__ void __launch_bounds__(CT_THREADS, 1)
qs_count_tiled_kernel(const CountTiledArgs a, const __grid_constant__ CUtensorMap tm16x16, const __grid_constant__ CUtensorMap tm8x16,
                      const __grid_constant__ CUtensorMap tm16x8) {
    // tmRxC: box of R rows x C cols
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* bufs = smem + 128;
    const int tid = threadIdx.x;
    constexpr uint32_t STAGE_BYTES = CT_TPS * CT_TREE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < CT_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    uint32_t* scratch = a.scratch + (size_t)blockIdx.x * (CT_TILE_Q * 3);

    // thread -> items (fixed for the whole kernel)
    const int x_ia = tid & 1, x_ib = (tid >> 1) & 1, x_c = (tid >> 2) & 15, x_d = tid >> 6;
    const int y_ia = tid & 1, y_b = (tid >> 1) & 15, y_c = tid >> 5;
    const uint32_t oXca = CT_OFF_CA + x_c * 32 + x_ia * 16, oXda = CT_OFF_DA + x_d * 32 + x_ia * 16;
    const uint32_t oXcb = CT_OFF_CB + x_c * 32 + x_ib * 16, oXdb = CT_OFF_DB + x_d * 32 + x_ib * 16;
    const uint32_t oYba = CT_OFF_BA + y_b * 32 + y_ia * 16, oYca = CT_OFF_CA + y_c * 32 + y_ia * 16;
    const uint32_t oYbd = CT_OFF_BD + y_b * 16, oYcd = CT_OFF_CD + y_c * 16;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const ushort4 tc = a.tiles[tile];
        const int a0 = tc.x * 16, b0 = tc.y * 16, c0 = tc.z * 16, d0 = tc.w * 8;

        XCounters xc_; YCounters yc_;
        zero(xc_); zero(yc_);

        const int nst = (a.m + CT_TPS - 1) / CT_TPS;
        // producer: lane l of warp 0 issues block (l % 7) of tree (l / 7) of a stage (7 * CT_TPS = 28 lanes)
        auto issue = [&](int s) {
            const int buf = s % CT_STAGES;
            const int nt = min(CT_TPS, a.m - s * CT_TPS);
            if (tid == 0) mbar_expect_tx(&full[buf], nt * CT_TREE_BYTES);
            __syncwarp();
            if (tid < 7 * nt) {
                const int tt = tid / 7, blk = tid % 7, tree = s * CT_TPS + tt;
                unsigned char* dst = bufs + buf * STAGE_BYTES + tt * CT_TREE_BYTES;
                switch (blk) {
                    case 0: tma_load_3d(dst + CT_OFF_CA, &tm16x16, a0, c0, tree, &full[buf]); break;
                    case 1: tma_load_3d(dst + CT_OFF_DA, &tm8x16, a0, d0, tree, &full[buf]); break;
                    case 2: tma_load_3d(dst + CT_OFF_CB, &tm16x16, b0, c0, tree, &full[buf]); break;
                    case 3: tma_load_3d(dst + CT_OFF_DB, &tm8x16, b0, d0, tree, &full[buf]); break;
                    case 4: tma_load_3d(dst + CT_OFF_BA, &tm16x16, a0, b0, tree, &full[buf]); break;
                    case 5: tma_load_3d(dst + CT_OFF_BD, &tm16x8, d0, b0, tree, &full[buf]); break;
                    default: tma_load_3d(dst + CT_OFF_CD, &tm16x8, d0, c0, tree, &full[buf]); break;
                }
            }
        };
        __syncthreads();   // previous tile's readers of the stage buffers and of the scratch are done
        if (tid < 32)
            for (int s = 0; s < CT_STAGES && s < nst; ++s) issue(s);

        int trees_in_chunk = 0;
        bool first_flush = true;
        auto flush = [&]() {
            // counters -> scratch (each (quartet, slot) has exactly one owner thread)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int bl = x_ib * 8 + j;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 v1 = __half22float2(xc_.s1[j][p]), v2 = __half22float2(xc_.s2[j][p]);
                    const int al = x_ia * 8 + 2 * p;
                    uint32_t* w = scratch + ((((x_d * 16 + x_c) * 16 + bl) * 16 + al) * 3);
                    if (first_flush) { w[1] = (uint32_t)v1.x; w[2] = (uint32_t)v2.x; w[4] = (uint32_t)v1.y; w[5] = (uint32_t)v2.y; }
                    else { w[1] += (uint32_t)v1.x; w[2] += (uint32_t)v2.x; w[4] += (uint32_t)v1.y; w[5] += (uint32_t)v2.y; }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 v0 = __half22float2(yc_.s0[j][p]);
                    const int al = y_ia * 8 + 2 * p;
                    uint32_t* w = scratch + ((((j * 16 + y_c) * 16 + y_b) * 16 + al) * 3);
                    if (first_flush) { w[0] = (uint32_t)v0.x; w[3] = (uint32_t)v0.y; }
                    else { w[0] += (uint32_t)v0.x; w[3] += (uint32_t)v0.y; }
                }
            }
            first_flush = false;
            zero(xc_); zero(yc_);
            trees_in_chunk = 0;
        };

        for (int s = 0; s < nst; ++s) {
            const int buf = s % CT_STAGES;
            mbar_wait(&full[buf], (phase >> buf) & 1u);
            phase ^= (1u << buf);
            const int nt = min(CT_TPS, a.m - s * CT_TPS);
            const unsigned char* base = bufs + buf * STAGE_BYTES;
#pragma unroll 1
            for (int tt = 0; tt < nt; ++tt, base += CT_TREE_BYTES) {
                role_x_step(xc_, lds128(base, oXca), lds128(base, oXda), lds128(base, oXcb), lds128(base, oXdb));
                role_y_step(yc_, lds128(base, oYba), lds128(base, oYca), lds128(base, oYbd), lds128(base, oYcd));
            }
            trees_in_chunk += nt;
            __syncthreads();
            if (tid < 32 && s + CT_STAGES < nst) issue(s + CT_STAGES);
            if (trees_in_chunk > 2048 - CT_TPS) flush();    // fp16 counters are exact up to 2048
        }
        flush();
        __syncthreads();

        // ---- epilogue: one half-warp per (b,c,d) triple, lane = a ----------------------------------
        if (a.table) {
            const int hw = tid >> 4, al = tid & 15;
            for (int trip = hw; trip < CT_TD * CT_TC * CT_TB; trip += CT_THREADS / 16) {
                const int bl = trip & 15, cl = (trip >> 4) & 15, dl = trip >> 8;
                const int av = a0 + al, bv = b0 + bl, cv = c0 + cl, dv = d0 + dl;
                if (!(av < bv && bv < cv && cv < dv && dv < a.d_end && dv >= a.d_begin)) continue;
                const uint32_t* w = scratch + (size_t)(trip * 16 + al) * 3;
                const uint64_t idx = quartet_rank(av, bv, cv, dv) - a.rank_base;
                switch (a.cint_bytes) {
                    case 1: ct_write_entry<uint8_t>(a.table, idx, w[0], w[1], w[2]); break;
                    case 2: ct_write_entry<uint16_t>(a.table, idx, w[0], w[1], w[2]); break;
                    case 4: ct_write_entry<uint32_t>(a.table, idx, w[0], w[1], w[2]); break;
                    default: ct_write_entry<unsigned long long>(a.table, idx, w[0], w[1], w[2]); break;
                }
            }
        }
        if (a.fused_score) {
            // thread = (b,c,d) triple, walks its 16 a's (kernels/score.cuh)
            for (int trip = tid; trip < CT_TD * CT_TC * CT_TB; trip += CT_THREADS) {
                const int bl = trip & 15, cl = (trip >> 4) & 15, dl = trip >> 8;
                const int bv = b0 + bl, cv = c0 + cl, dv = d0 + dl;
                if (!(bv < cv && cv < dv && dv < a.d_end && dv >= a.d_begin)) continue;
                const ScoreArgs& sa = a.sa;
                const int q = sa.lca[(size_t)bv * sa.n + cv], r = sa.lca[(size_t)cv * sa.n + dv];
                const int dq = sa.idepth[q], dr = sa.idepth[r];
                const uint16_t* lrow = sa.lca + (size_t)bv * sa.n;
                PairAcc acc; acc.s1 = acc.s2 = acc.s3 = 0; acc.best = QS_TRIPLE_NONE; acc.best_q = INFINITY; acc.key = -1;
                unsigned long long memo_t = QS_TRIPLE_NONE; double memo_q = 0.0;
                int last_p = -1, key = -1, rslot = -1;
                const int amax = min(16, bv - a0);
                for (int al = 0; al < amax; ++al) {
                    const int p = lrow[a0 + al];
                    if (p != last_p) { last_p = p; key = quartet_pair_key(sa, p, q, r, sa.idepth[p], dq, dr, rslot); }
                    if (key < 0) continue;
                    const uint32_t* w = scratch + (size_t)(trip * 16 + al) * 3;
                    const unsigned long long k0 = ((unsigned long long)w[0] * sa.count_scale) & sa.cint_mask;
                    const unsigned long long k1 = ((unsigned long long)w[1] * sa.count_scale) & sa.cint_mask;
                    const unsigned long long k2 = ((unsigned long long)w[2] * sa.count_scale) & sa.cint_mask;
                    pair_add(sa, acc, key, rslot, k0, k1, k2, memo_t, memo_q);
                }
                pair_flush(sa, acc);
            }
        }
    }
}

}  // namespace qs
