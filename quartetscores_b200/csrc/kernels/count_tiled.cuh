// count_tiled.cuh — quartet counting kernel for any n: Cartesian 4-D tiles of the quartet space with the
// needed sub-blocks of every gene tree's distance matrix staged in shared memory by TMA tensor loads.
//
// Same arithmetic as count_items.cuh (fixed-pair four-point test, HSET2 masks + three-input IADD3,
// counters in registers across all trees, class-A trees skip role Y).  A tile is A x B x C x D =
// 16 x 16 x 16 x 8 taxon ids (32,768 quartets).  Its 512 threads run two phases per tile:
//   phase X: thread = one (c,d) pair x 8 a x 8 b; needs the sub-blocks CxA, DxA, CxB, DxB of D   (all trees)
//   phase Y: thread = one (b,c) pair x 8 a x 8 d; needs BxA, CxA, BxD, CxD                       (class-B trees)
// = 1,536 bytes per tree and phase for 32,768 evaluations (0.05 B/eval), fetched with
// cp.async.bulk.tensor.3d (SASS UTMALDG) from the [tree][row][col] tensor, one tree PAIR per pipeline
// stage; an odd tree count is padded with an out-of-bounds tree coordinate, which TMA zero-fills
// (an all-zero matrix never satisfies a strict inequality, so it counts nothing).
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
constexpr int CT_TREE_BYTES = 1536;                              // per tree and phase
constexpr int CT_STAGE_BYTES = 2 * CT_TREE_BYTES;                // one tree pair
constexpr int CT_STAGES = 8;
// block offsets inside one tree's record
constexpr int CT_X_CA = 0, CT_X_DA = 512, CT_X_CB = 768, CT_X_DB = 1280;
constexpr int CT_Y_BA = 0, CT_Y_CA = 512, CT_Y_BD = 1024, CT_Y_CD = 1280;

struct CountTiledArgs {
    const ushort4* tiles;       // (iA, iB, iC, jD): A/B/C ranges start at 16*i, D range at d_tile_base + 8*j
    int d_tile_base;            // first d of this shard
    int n_tiles;
    int n, m;
    int d_begin, d_end;
    const int32_t* order;       // [m] class-sorted tree order (class A first)
    const int32_t* n_class_a;   // device scalar |A|
    int* tile_counter;          // zeroed by the caller: dynamic tile scheduling
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
    __shared__ int s_tile;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    int* done = reinterpret_cast<int*>(smem + 64);
    unsigned char* bufs = smem + 128;
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < CT_STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    uint32_t phase = 0;
    uint32_t* scratch = a.scratch + (size_t)blockIdx.x * (CT_TILE_Q * 3);
    const int mA = *a.n_class_a;

    // thread -> items (fixed for the whole kernel)
    const int x_ia = tid & 1, x_ib = (tid >> 1) & 1, x_c = (tid >> 2) & 15, x_d = tid >> 6;
    const int y_ia = tid & 1, y_b = (tid >> 1) & 15, y_c = tid >> 5;
    const uint32_t oXca = CT_X_CA + x_c * 32 + x_ia * 16, oXda = CT_X_DA + x_d * 32 + x_ia * 16;
    const uint32_t oXcb = CT_X_CB + x_c * 32 + x_ib * 16, oXdb = CT_X_DB + x_d * 32 + x_ib * 16;
    const uint32_t oYba = CT_Y_BA + y_b * 32 + y_ia * 16, oYca = CT_Y_CA + y_c * 32 + y_ia * 16;
    const uint32_t oYbd = CT_Y_BD + y_b * 16, oYcd = CT_Y_CD + y_c * 16;

    while (true) {
        __syncthreads();     // previous tile: scratch readers and stage readers are done
        if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= a.n_tiles) break;
        const ushort4 tc = a.tiles[tile];
        const int a0 = tc.x * 16, b0 = tc.y * 16, c0 = tc.z * 16, d0 = a.d_tile_base + tc.w * 8;

        // stream the tree pairs [t0,t1) of the class-sorted order; PHASE_Y selects the four sub-blocks
        auto stream = [&](int t0, int t1, bool phase_y, auto&& f) {
            const int npairs = (t1 - t0 + 1) >> 1;
            auto issue = [&](int pr, int tree_a, int tree_b) {   // one thread
                const int buf = pr % CT_STAGES;
                mbar_expect_tx(&full[buf], CT_STAGE_BYTES);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tree = e ? tree_b : tree_a;
                    unsigned char* dst = bufs + buf * CT_STAGE_BYTES + e * CT_TREE_BYTES;
                    if (!phase_y) {
                        tma_load_3d(dst + CT_X_CA, &tm16x16, a0, c0, tree, &full[buf]);
                        tma_load_3d(dst + CT_X_DA, &tm8x16, a0, d0, tree, &full[buf]);
                        tma_load_3d(dst + CT_X_CB, &tm16x16, b0, c0, tree, &full[buf]);
                        tma_load_3d(dst + CT_X_DB, &tm8x16, b0, d0, tree, &full[buf]);
                    } else {
                        tma_load_3d(dst + CT_Y_BA, &tm16x16, a0, b0, tree, &full[buf]);
                        tma_load_3d(dst + CT_Y_CA, &tm16x16, a0, c0, tree, &full[buf]);
                        tma_load_3d(dst + CT_Y_BD, &tm16x8, d0, b0, tree, &full[buf]);
                        tma_load_3d(dst + CT_Y_CD, &tm16x8, d0, c0, tree, &full[buf]);
                    }
                }
            };
            auto tree_of = [&](int t) -> int { return t < t1 ? a.order[t] : a.m; };   // out of bounds -> TMA zero fill
            __syncthreads();             // every warp is done with every stage of the previous stream
            if (tid < CT_STAGES) {
                done[tid] = 0;
                if (tid < npairs) issue(tid, tree_of(t0 + 2 * tid), tree_of(t0 + 2 * tid + 1));
            }
            __syncthreads();
            for (int pr = 0; pr < npairs; ++pr) {
                const int buf = pr % CT_STAGES;
                // every warp prefetches the tree ids of the pair that will refill this stage (see count_items.cuh)
                int nxt_a = 0, nxt_b = 0;
                const int prn = pr + CT_STAGES;
                if (lane == 0 && prn < npairs) { nxt_a = tree_of(t0 + 2 * prn); nxt_b = tree_of(t0 + 2 * prn + 1); }
                mbar_wait(&full[buf], (phase >> buf) & 1u);
                phase ^= (1u << buf);
                const unsigned char* base = bufs + buf * CT_STAGE_BYTES;
                f(base, base + CT_TREE_BYTES);
                __syncwarp();
                if (lane == 0) {
                    const int old = atomicAdd(&done[buf], 1);
                    if (old == CT_THREADS / 32 - 1) {
                        done[buf] = 0;
                        if (prn < npairs) issue(prn, nxt_a, nxt_b);
                    }
                }
            }
        };

        bool first_flush = true;
        // ---- phase X: slots 1 and 2 over all trees; class-A trees also determine slot 0 ----------------
        {
            XCounters x;
            auto flush_x = [&](bool sub0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int bl = x_ib * 8 + j;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t g0, g1, l0, l1;
                        decode(x.gt[j][p], g0, g1);
                        decode(x.lt[j][p], l0, l1);
                        const int al = x_ia * 8 + 2 * p;
                        uint32_t* w = scratch + ((((x_d * 16 + x_c) * 16 + bl) * 16 + al) * 3);
                        const uint32_t z0 = sub0 ? 0u - (g0 + l0) : 0u, z1 = sub0 ? 0u - (g1 + l1) : 0u;
                        if (first_flush) { w[0] = z0; w[1] = g0; w[2] = l0; w[3] = z1; w[4] = g1; w[5] = l1; }
                        else { w[0] += z0; w[1] += g0; w[2] += l0; w[3] += z1; w[4] += g1; w[5] += l1; }
                    }
                }
                first_flush = false;
            };
            auto step = [&](const unsigned char* s0, const unsigned char* s1) {
                BlockRows r0{lds128(s0, oXca), lds128(s0, oXda), lds128(s0, oXcb), lds128(s0, oXdb)};
                BlockRows r1{lds128(s1, oXca), lds128(s1, oXda), lds128(s1, oXcb), lds128(s1, oXdb)};
                step_gt_lt(x, r0, r1);
            };
            for (int cls = 0; cls < 2; ++cls) {      // m > 0, so at least one chunk runs and initialises the scratch
                const int lo = cls == 0 ? 0 : mA, hi = cls == 0 ? mA : a.m;
                for (int t0 = lo; t0 < hi; t0 += QS_MAX_CHUNK_TREES) {
                    zero(x);
                    stream(t0, min(hi, t0 + QS_MAX_CHUNK_TREES), false, step);
                    flush_x(cls == 0);
                }
            }
        }
        // ---- phase Y: slot 0 over the class-B trees ------------------------------------------------------
        if (mA < a.m) {
            GCounters y;
            auto step = [&](const unsigned char* s0, const unsigned char* s1) {
                BlockRows r0{lds128(s0, oYba), lds128(s0, oYca), lds128(s0, oYbd), lds128(s0, oYcd)};
                BlockRows r1{lds128(s1, oYba), lds128(s1, oYca), lds128(s1, oYbd), lds128(s1, oYcd)};
                step_gt(y, r0, r1);
            };
            __syncthreads();             // phase-X flushes of other threads to the same scratch entries are visible
            for (int t0 = mA; t0 < a.m; t0 += QS_MAX_CHUNK_TREES) {
                const int t1 = min(a.m, t0 + QS_MAX_CHUNK_TREES);
                zero(y);
                stream(t0, t1, true, step);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h0, h1;
                        decode(y.gt[j][p], h0, h1);
                        const int al = y_ia * 8 + 2 * p;
                        uint32_t* w = scratch + ((((j * 16 + y_c) * 16 + y_b) * 16 + al) * 3);
                        w[0] += h0; w[3] += h1;
                    }
                }
            }
        }
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
                const uint32_t w0 = w[0] + (uint32_t)mA;
                switch (a.cint_bytes) {
                    case 1: ct_write_entry<uint8_t>(a.table, idx, w0, w[1], w[2]); break;
                    case 2: ct_write_entry<uint16_t>(a.table, idx, w0, w[1], w[2]); break;
                    case 4: ct_write_entry<uint32_t>(a.table, idx, w0, w[1], w[2]); break;
                    default: ct_write_entry<unsigned long long>(a.table, idx, w0, w[1], w[2]); break;
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
                    const unsigned long long k0 = ((unsigned long long)(w[0] + (uint32_t)mA) * sa.count_scale) & sa.cint_mask;
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
