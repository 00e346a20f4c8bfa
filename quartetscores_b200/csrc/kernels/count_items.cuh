// count_items.cuh — quartet counting kernel for reference trees whose whole n x n distance matrix fits in
// shared memory at least four times (n <= ~165): BASELINE config 1/2 shapes.
//
// Replaces the reference hot loop QuartetCounterLookup::updateQuartetsThreeClades
// (src/QuartetCounterLookup.hpp:66-106) + countQuartets (:197-238).  Instead of enumerating clades per
// tree and scattering increments into an n^4 table, every thread OWNS a fixed set of quartets, keeps
// their topology counters in registers across all gene trees of its chunk, and decides each
// (quartet, tree) with the four-point condition on the tree's distance matrix D, which is staged in
// shared memory by TMA bulk copies (cp.async.bulk, SASS UBLKCP) through an mbarrier pipeline.
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
// Fully resolved trees need no role Y.  The distance kernel classifies every gene tree: class A = all n
// taxa present and no node of degree > 3, i.e. every quartet is resolved in it, so
// slot0 + slot1 + slot2 = 1 per tree and slot 0 = |A| - slot1_A - slot2_A.  Class-A trees (first in the
// class-sorted order[]) are only run through role X (2 compares per quartet x tree instead of 3); their
// slot-1/2 hits are accumulated apart and the narrowing kernel derives slot 0.
//
// Work items (host-built tables, ContextItems in qscuda.cu), all 8 x 8 register blocks:
//   XO  (c,d) x a-block ia x b-block ib, ia < ib : G(a)>G(b) and G(a)<G(b)          128 compares / tree pair
//   XD  (c,d) x diagonal block i (a and b in the same block): all ordered pairs, G(x)>G(y) only;
//       x<y gives slot 1 of (x,y), x>y gives slot 2 of (y,x)                        2 items / thread
//   Y   (b,c) x a-block x d-block : G(a)>G(d)                                       2 items / thread
// so a thread always carries 64 counter registers and issues 128 HSET2 + 64 IADD3 per tree pair.
// Tasks = (kind, 512 thread-items, tree chunk) are handed to persistent CTAs through an atomic counter.
#pragma once
#include "common.cuh"
#include "count_roles.cuh"

namespace qs {

struct CountItem { uint16_t p, q, iu, iv; };     // fixed pair (p,q); u-block and v-block index (blocks of 8 taxa)

enum { ITEM_XO = 0, ITEM_XD = 1, ITEM_Y = 2 };

struct CountTask {
    int32_t kind;        // ITEM_*
    int32_t first;       // first item of this task in the kind's item array
    int32_t count;       // number of items (<= THREADS * items-per-thread)
    int32_t cls;         // 0 = class-A trees (fully resolved, complete), 1 = class-B trees
    int32_t chunk, nchunks;
};

struct CountItemsArgs {
    const __half* D;            // [m][n][n_pad] fp16, NaN = missing, indexed by ORIGINAL tree index
    const int32_t* order;       // [m] class-sorted tree order: class A first
    const int32_t* n_class_a;   // device scalar |A|
    const __half* nan_tree;     // one all-NaN matrix (pads an odd tree count to a pair)
    const CountItem* items[3];
    const CountTask* tasks;
    int n_tasks;
    int* task_counter;          // zeroed by the caller
    uint32_t* ws;               // [(rank - rank_base)*QS_WS_SLOTS + k] uint32 workspace (zeroed by caller), see ci_add
    uint64_t rank_base;
    int n, n_pad, m;
    int d_begin, d_end;         // shard: quartets with d in [d_begin, d_end)
    int n_stages;               // pipeline depth, one tree pair per stage
    uint32_t tree_bytes;        // n * n_pad * 2
};

constexpr int CI_MAX_STAGES = 8;

struct ItemPipe {
    uint64_t* full;             // [CI_MAX_STAGES] tx barriers
    int* done;                  // [CI_MAX_STAGES] warps finished with the stage
    unsigned char* bufs;
    uint32_t stage_bytes;
    uint32_t phase;             // bit s = parity to wait for on full[s]
};

// Stream the tree pairs [t0,t1) of the class-sorted order through the pipeline; f(base0, base1) is called
// once per pair with the shared-memory addresses of the two matrices.  Warps run independently: the
// last warp to finish a stage refills it (no CTA-wide barrier inside the loop).
template <int THREADS, class F>
__device__ __forceinline__ void stream_tree_pairs(const CountItemsArgs& a, ItemPipe& P, int t0, int t1, F&& f) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int npairs = (t1 - t0 + 1) >> 1;
    const size_t tree_elems = a.tree_bytes >> 1;
    auto issue = [&](int pr, int tree_a, int tree_b) {       // one thread; tree_b < 0: pad the pair with the NaN matrix
        const int buf = pr % a.n_stages;
        unsigned char* dst = P.bufs + (size_t)buf * P.stage_bytes;
        mbar_expect_tx(&P.full[buf], 2 * a.tree_bytes);
        bulk_g2s(dst, a.D + (size_t)tree_a * tree_elems, a.tree_bytes, &P.full[buf]);
        bulk_g2s(dst + a.tree_bytes, tree_b >= 0 ? a.D + (size_t)tree_b * tree_elems : a.nan_tree, a.tree_bytes, &P.full[buf]);
    };
    auto tree_of = [&](int t) -> int { return t < t1 ? a.order[t] : -1; };
    __syncthreads();                 // the previous task's readers are done with every stage
    if (tid < a.n_stages) {
        P.done[tid] = 0;
        if (tid < npairs) issue(tid, tree_of(t0 + 2 * tid), tree_of(t0 + 2 * tid + 1));
    }
    __syncthreads();
    for (int pr = 0; pr < npairs; ++pr) {
        const int buf = pr % a.n_stages;
        // every warp prefetches the tree ids of the pair that will refill this stage, so that whichever warp
        // leaves the stage last can issue the copies without waiting on global memory
        int nxt_a = -1, nxt_b = -1;
        const int prn = pr + a.n_stages;
        if (lane == 0 && prn < npairs) { nxt_a = tree_of(t0 + 2 * prn); nxt_b = tree_of(t0 + 2 * prn + 1); }
        mbar_wait(&P.full[buf], (P.phase >> buf) & 1u);
        P.phase ^= (1u << buf);
        const unsigned char* base = P.bufs + (size_t)buf * P.stage_bytes;
        f(base, base + a.tree_bytes);
        __syncwarp();
        if (lane == 0) {
            const int old = atomicAdd(&P.done[buf], 1);
            if (old == THREADS / 32 - 1) {           // last warp out refills the stage
                P.done[buf] = 0;
                if (prn < npairs) issue(prn, nxt_a, nxt_b);
            }
        }
    }
}

__device__ __forceinline__ void ci_tree_range(const CountItemsArgs& a, const CountTask& T, int& t0, int& t1) {
    const int mA = *a.n_class_a;
    const int lo = T.cls == 0 ? 0 : mA, hi = T.cls == 0 ? mA : a.m;
    int per = (hi - lo + T.nchunks - 1) / T.nchunks;
    per = (per + 1) & ~1;                            // whole pairs
    t0 = lo + T.chunk * per;
    t1 = min(hi, t0 + per);
}

// Workspace entry = QS_WS_SLOTS uint32: [slot0_B, slot1_B, slot2_B, slot1_A, slot2_A] — class-A role-X hits are kept
// apart so that the narrowing kernel can derive their slot 0 (|A| - slot1_A - slot2_A) without extra atomics here.
constexpr int QS_WS_SLOTS = 5;
__device__ __forceinline__ void ci_add(uint32_t* w, int slot, uint32_t v, bool cls_a) {
    if (v) atomicAdd(w + slot + (cls_a ? 2 : 0), v);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) qs_count_items_kernel(const CountItemsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_task;
    ItemPipe P;
    P.full = reinterpret_cast<uint64_t*>(smem);
    P.done = reinterpret_cast<int*>(smem + 64);
    P.bufs = smem + 128;
    P.stage_bytes = 2 * a.tree_bytes;
    P.phase = 0;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < CI_MAX_STAGES; ++s) mbar_init(&P.full[s], 1);
        fence_mbar_init();
    }
    const uint32_t rowb = (uint32_t)a.n_pad * 2u;

    while (true) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(a.task_counter, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= a.n_tasks) break;
        const CountTask T = a.tasks[task];
        int t0, t1;
        ci_tree_range(a, T, t0, t1);
        if (t0 >= t1) continue;
        const bool sub0 = (T.cls == 0);

        if (T.kind == ITEM_XO) {
            const bool valid = tid < T.count;
            CountItem it = valid ? a.items[ITEM_XO][T.first + tid] : CountItem{0, 0, 0, 0};
            const uint32_t oPu = it.p * rowb + it.iu * 16u, oQu = it.q * rowb + it.iu * 16u;
            const uint32_t oPv = it.p * rowb + it.iv * 16u, oQv = it.q * rowb + it.iv * 16u;
            XCounters x; zero(x);
            stream_tree_pairs<THREADS>(a, P, t0, t1, [&](const unsigned char* b0, const unsigned char* b1) {
                BlockRows r0{lds128(b0, oPu), lds128(b0, oQu), lds128(b0, oPv), lds128(b0, oQv)};
                BlockRows r1{lds128(b1, oPu), lds128(b1, oQu), lds128(b1, oPv), lds128(b1, oQv)};
                step_gt_lt(x, r0, r1);
            });
            if (valid) {
                const int c = it.p, d = it.q;
                const uint64_t rcd = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int b = it.iv * 8 + j;
                    if (b >= c) continue;
                    uint32_t* wb = a.ws + (rcd + (uint64_t)b * (b - 1) / 2 + it.iu * 8) * QS_WS_SLOTS;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t g0, g1, l0, l1;
                        decode(x.gt[j][p], g0, g1);
                        decode(x.lt[j][p], l0, l1);
                        uint32_t* w = wb + (2 * p) * QS_WS_SLOTS;
                        ci_add(w, 1, g0, sub0); ci_add(w, 2, l0, sub0);
                        ci_add(w + QS_WS_SLOTS, 1, g1, sub0); ci_add(w + QS_WS_SLOTS, 2, l1, sub0);
                    }
                }
            }
        } else if (T.kind == ITEM_XD) {
            const int i0 = T.first + 2 * tid, i1 = i0 + 1, iend = T.first + T.count;
            const bool v0 = i0 < iend, v1 = i1 < iend;
            CountItem A = v0 ? a.items[ITEM_XD][i0] : CountItem{0, 0, 0, 0};
            CountItem B = v1 ? a.items[ITEM_XD][i1] : A;
            const uint32_t oAp = A.p * rowb + A.iu * 16u, oAq = A.q * rowb + A.iu * 16u;
            const uint32_t oBp = B.p * rowb + B.iu * 16u, oBq = B.q * rowb + B.iu * 16u;
            GCounters ga, gb; zero(ga); zero(gb);
            stream_tree_pairs<THREADS>(a, P, t0, t1, [&](const unsigned char* b0, const unsigned char* b1) {
                step_gt_diag(ga, lds128(b0, oAp), lds128(b0, oAq), lds128(b1, oAp), lds128(b1, oAq));
                step_gt_diag(gb, lds128(b0, oBp), lds128(b0, oBq), lds128(b1, oBp), lds128(b1, oBq));
            });
            auto flush = [&](const CountItem& it, const GCounters& g) {
                const int c = it.p, d = it.q, x0 = it.iu * 8;
                const uint64_t rcd = binom4((uint64_t)d) + binom3((uint64_t)c) - a.rank_base;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int y = x0 + j;                          // the "v" taxon
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h[2];
                        decode(g.gt[j][p], h[0], h[1]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int x = x0 + 2 * p + e;          // the "u" taxon: counted G(x) > G(y)
                            if (x == y || max(x, y) >= c) continue;
                            const int lo = min(x, y), hi = max(x, y);
                            uint32_t* w = a.ws + (rcd + (uint64_t)hi * (hi - 1) / 2 + lo) * QS_WS_SLOTS;
                            ci_add(w, (x < y) ? 1 : 2, h[e], sub0);   // x<y: G(a)>G(b) slot 1;  x>y: G(b)>G(a) slot 2
                        }
                    }
                }
            };
            if (v0) flush(A, ga);
            if (v1) flush(B, gb);
        } else {   // ITEM_Y
            const int i0 = T.first + 2 * tid, i1 = i0 + 1, iend = T.first + T.count;
            const bool v0 = i0 < iend, v1 = i1 < iend;
            CountItem A = v0 ? a.items[ITEM_Y][i0] : CountItem{0, 0, 0, 0};
            CountItem B = v1 ? a.items[ITEM_Y][i1] : A;
            const uint32_t oApu = A.p * rowb + A.iu * 16u, oAqu = A.q * rowb + A.iu * 16u, oApv = A.p * rowb + A.iv * 16u, oAqv = A.q * rowb + A.iv * 16u;
            const uint32_t oBpu = B.p * rowb + B.iu * 16u, oBqu = B.q * rowb + B.iu * 16u, oBpv = B.p * rowb + B.iv * 16u, oBqv = B.q * rowb + B.iv * 16u;
            GCounters ga, gb; zero(ga); zero(gb);
            stream_tree_pairs<THREADS>(a, P, t0, t1, [&](const unsigned char* b0, const unsigned char* b1) {
                {
                    BlockRows r0{lds128(b0, oApu), lds128(b0, oAqu), lds128(b0, oApv), lds128(b0, oAqv)};
                    BlockRows r1{lds128(b1, oApu), lds128(b1, oAqu), lds128(b1, oApv), lds128(b1, oAqv)};
                    step_gt(ga, r0, r1);
                }
                {
                    BlockRows r0{lds128(b0, oBpu), lds128(b0, oBqu), lds128(b0, oBpv), lds128(b0, oBqv)};
                    BlockRows r1{lds128(b1, oBpu), lds128(b1, oBqu), lds128(b1, oBpv), lds128(b1, oBqv)};
                    step_gt(gb, r0, r1);
                }
            });
            auto flush = [&](const CountItem& it, const GCounters& g) {
                const int b = it.p, c = it.q;
                const int dlo = max(c + 1, a.d_begin);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int d = it.iv * 8 + j;
                    if (d < dlo || d >= a.d_end) continue;
                    uint32_t* wb = a.ws + (binom4((uint64_t)d) + binom3((uint64_t)c) + (uint64_t)b * (b - 1) / 2 - a.rank_base + it.iu * 8) * QS_WS_SLOTS;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        uint32_t h0, h1;
                        decode(g.gt[j][p], h0, h1);
                        const int a0 = it.iu * 8 + 2 * p;
                        if (a0 < b && h0) atomicAdd(wb + (2 * p) * QS_WS_SLOTS, h0);
                        if (a0 + 1 < b && h1) atomicAdd(wb + (2 * p + 1) * QS_WS_SLOTS, h1);
                    }
                }
            };
            if (v0) flush(A, ga);
            if (v1) flush(B, gb);
        }
    }
}

// uint32 workspace -> CINT table (QuartetLookupTable layout); thread = one table entry
template <typename CINT>
__global__ void qs_narrow_kernel(const uint32_t* __restrict__ ws, CINT* __restrict__ table, uint64_t n_entries, const int32_t* __restrict__ n_class_a) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t mA = (uint32_t)*n_class_a;
    for (; i < n_entries; i += stride) {
        const uint32_t* w = ws + i * QS_WS_SLOTS;
        const uint32_t s1a = w[3], s2a = w[4];
        table[i * 3 + 0] = (CINT)(w[0] + mA - s1a - s2a);
        table[i * 3 + 1] = (CINT)(w[1] + s1a);
        table[i * 3 + 2] = (CINT)(w[2] + s2a);
    }
}

}  // namespace qs
