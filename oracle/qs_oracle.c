/* oracle/qs_oracle.c — TEST INFRASTRUCTURE ONLY (see qs_oracle.h).
 *
 * Plain-C restatement of the reference's quartet counting and scoring.  Written from the
 * behaviour of /root/reference/src (file:line cited per function); no reference code is copied.
 * Scalar, simple and slow on purpose: it is the checker, never the thing measured or shipped.
 */
#include "qs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* table layout: quartet_lookup_table.hpp                                                      */
/* ------------------------------------------------------------------------------------------ */

uint64_t qso_num_quartets(uint64_t n) { return n < 4 ? 0 : n * (n - 1) * (n - 2) * (n - 3) / 24; }

/* quartet_lookup_table.hpp:141-168: C(s3,4)+C(s2,3)+C(s1,2)+s0 for s3>s2>s1>s0 */
static uint64_t rank_sorted(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3) {
    return s3 * (s3 - 1) * (s3 - 2) * (s3 - 3) / 24 + s2 * (s2 - 1) * (s2 - 2) / 6 + s1 * (s1 - 1) / 2 + s0;
}

/* quartet_lookup_table.hpp:170-212 (sorting network on the two pairs) */
uint64_t qso_rank(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    uint64_t v[4] = {a, b, c, d};
    for (int i = 1; i < 4; ++i) {
        uint64_t x = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
        v[j + 1] = x;
    }
    return rank_sorted(v[0], v[1], v[2], v[3]);
}

/* quartet_lookup_table.hpp:87-111: slot 0 = {s0,s1}|{s2,s3}, 1 = {s0,s2}|{s1,s3}, 2 = {s0,s3}|{s1,s2} */
int qso_tuple_index(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    int ac = a < c, ad = a < d, bc = b < c, bd = b < d;
    int x = (ac & ad & bc & bd) | (!ac & !bc & !ad & !bd);           /* pairs do not interleave */
    int ab_in_cd = (!ac & ad & !bc & bd) | (!ad & ac & !bd & bc);    /* one pair nested in the other */
    int cd_in_ab = (ac & !bc & ad & !bd) | (bc & !ac & bd & !ad);
    int z = ab_in_cd | cd_in_ab;
    int y = !x & !z;
    return y + 2 * z;
}

/* ------------------------------------------------------------------------------------------ */
/* small rooted-tree helper built from the flat encoding                                       */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int n;              /* nodes */
    const int32_t* parent;
    const int32_t* leaf_id;
    int* depth;         /* TreeInformation.hpp:96 dist_to_root */
    int* first_child;   /* children in sibling order */
    int* next_sib;
    int* n_child;
    int* lo;            /* subtree = leaves [lo,hi) of the Euler-tour leaf sequence */
    int* hi;
    int* leaf_seq;      /* node index of k-th leaf in Euler-tour order */
    int k;              /* number of leaves */
} otree;

static void otree_free(otree* t) {
    free(t->depth); free(t->first_child); free(t->next_sib); free(t->n_child); free(t->lo); free(t->hi); free(t->leaf_seq);
}

/* child_rank may be NULL: siblings then ordered by node index. */
static int otree_build(otree* t, int n, const int32_t* parent, const int32_t* leaf_id, const int32_t* child_rank) {
    memset(t, 0, sizeof(*t));
    t->n = n; t->parent = parent; t->leaf_id = leaf_id;
    t->depth = (int*)calloc(n, sizeof(int)); t->first_child = (int*)malloc(n * sizeof(int));
    t->next_sib = (int*)malloc(n * sizeof(int)); t->n_child = (int*)calloc(n, sizeof(int));
    t->lo = (int*)malloc(n * sizeof(int)); t->hi = (int*)malloc(n * sizeof(int)); t->leaf_seq = (int*)malloc(n * sizeof(int));
    for (int i = 0; i < n; ++i) { t->first_child[i] = -1; t->next_sib[i] = -1; }
    if (n == 0 || parent[0] != -1) return -1;
    for (int i = 1; i < n; ++i) {
        if (parent[i] < 0 || parent[i] >= i) return -1;
        t->depth[i] = t->depth[parent[i]] + 1;
        t->n_child[parent[i]]++;
    }
    /* insert children sorted by (child_rank, index): walk indices downwards and push to front, then
       stable insertion by rank */
    for (int i = n - 1; i >= 1; --i) {
        int p = parent[i];
        if (!child_rank) { t->next_sib[i] = t->first_child[p]; t->first_child[p] = i; continue; }
        int prev = -1, cur = t->first_child[p];
        while (cur != -1 && (child_rank[cur] < child_rank[i] || (child_rank[cur] == child_rank[i] && cur < i))) { prev = cur; cur = t->next_sib[cur]; }
        t->next_sib[i] = cur;
        if (prev == -1) t->first_child[p] = i; else t->next_sib[prev] = i;
    }
    /* iterative DFS (Euler tour, genesis eulertour.hpp:114-121) -> leaf sequence and [lo,hi) */
    int* stack = (int*)malloc((n + 1) * sizeof(int));
    int* it = (int*)malloc(n * sizeof(int));
    int sp = 0; t->k = 0;
    stack[sp++] = 0; it[0] = t->first_child[0]; t->lo[0] = 0;
    if (t->first_child[0] == -1) { t->leaf_seq[t->k++] = 0; }
    while (sp > 0) {
        int v = stack[sp - 1];
        int c = it[v];
        if (c == -1) { t->hi[v] = t->k; --sp; continue; }
        it[v] = t->next_sib[c];
        t->lo[c] = t->k;
        if (t->first_child[c] == -1) { t->leaf_seq[t->k++] = c; t->hi[c] = t->k; }
        else { stack[sp++] = c; it[c] = t->first_child[c]; }
    }
    free(stack); free(it);
    return 0;
}

/* TreeInformation.hpp:71-76 with the tree's own root (Euler tour + RMQ there; plain climbing here) */
static int lca0(const otree* t, int u, int v) {
    while (t->depth[u] > t->depth[v]) u = t->parent[u];
    while (t->depth[v] > t->depth[u]) v = t->parent[v];
    while (u != v) { u = t->parent[u]; v = t->parent[v]; }
    return u;
}
/* TreeInformation.hpp:71-90: LCA with respect to another root, "odd man out" */
static int lca_r(const otree* t, int u, int v, int r) {
    if (r == 0) return lca0(t, u, v);
    int c1 = lca0(t, u, v), c2 = lca0(t, u, r), c3 = lca0(t, v, r);
    if (c1 == c2) return c3;
    if (c1 == c3) return c2;
    return c1;
}
/* TreeInformation.hpp:40-43 */
static unsigned dist_edges(const otree* t, int u, int v) {
    int l = lca0(t, u, v);
    return (unsigned)(t->depth[u] + t->depth[v] - 2 * t->depth[l]);
}

/* ------------------------------------------------------------------------------------------ */
/* distances                                                                                   */
/* ------------------------------------------------------------------------------------------ */

int qso_distance_matrix(int n_nodes, const int32_t* parent, const int32_t* leaf_id, int n_taxa, uint16_t* D) {
    otree t;
    if (otree_build(&t, n_nodes, parent, leaf_id, NULL) != 0) { otree_free(&t); return -1; }
    for (size_t i = 0; i < (size_t)n_taxa * n_taxa; ++i) D[i] = QSO_MISSING;
    int mx = 0;
    for (int i = 0; i < t.k; ++i) {
        int u = t.leaf_seq[i], x = leaf_id[u];
        if (x < 0 || x >= n_taxa) { otree_free(&t); return -2; }
        for (int j = 0; j < t.k; ++j) {
            int v = t.leaf_seq[j], y = leaf_id[v];
            if (y < 0 || y >= n_taxa) { otree_free(&t); return -2; }
            unsigned d = dist_edges(&t, u, v);
            if ((int)d > mx) mx = (int)d;
            D[(size_t)x * n_taxa + y] = (uint16_t)d;
        }
    }
    otree_free(&t);
    return mx;
}

/* ------------------------------------------------------------------------------------------ */
/* counting: clade enumeration (QuartetCounterLookup.hpp)                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int start, end; } crange; /* cyclic [start,end) over the k Euler-tour leaves */

typedef struct {
    int n; uint64_t n2, n3;
    uint32_t* fast;     /* n^4 or NULL */
    uint32_t* compact;  /* C(n,4)*3 or NULL */
} ctable;

/* QuartetCounterLookup.hpp:66-106 */
static void three_clades(ctable* T, crange s1, crange s2, crange s3, const int* el, int k) {
    for (int ai = s1.start; ai != s1.end; ai = (ai + 1) % k) {
        uint64_t a = (uint64_t)el[ai];
        for (int a2i = (ai + 1) % k; a2i != s1.end; a2i = (a2i + 1) % k) {
            uint64_t a2 = (uint64_t)el[a2i];
            for (int bi = s2.start; bi != s2.end; bi = (bi + 1) % k) {
                uint64_t b = (uint64_t)el[bi];
                for (int ci = s3.start; ci != s3.end; ci = (ci + 1) % k) {
                    uint64_t c = (uint64_t)el[ci];
                    if (T->compact) T->compact[qso_rank(a, a2, b, c) * 3 + qso_tuple_index(a, a2, b, c)]++;  /* :83-87 */
                    else T->fast[a * T->n3 + a2 * T->n2 + b * T->n + c]++;                                   /* :90   */
                }
            }
        }
    }
}

/* QuartetCounterLookup.hpp:197-238 for one tree (Euler leaf list :211-221, per node :223-228,
 * link triples :167-188, three role assignments :135-154, cyclic leaf ranges :117-121) */
static int count_tree_clades(ctable* T, int n_nodes, const int32_t* parent, const int32_t* leaf_id) {
    otree t;
    if (otree_build(&t, n_nodes, parent, leaf_id, NULL) != 0) { otree_free(&t); return -1; }
    int k = t.k;
    int* el = (int*)malloc((k > 0 ? k : 1) * sizeof(int));
    for (int i = 0; i < k; ++i) {
        el[i] = leaf_id[t.leaf_seq[i]];
        if (el[i] < 0 || el[i] >= T->n) { free(el); otree_free(&t); return -2; }  /* :218 throws on unknown taxon */
    }
    crange* links = (crange*)malloc((n_nodes + 1) * sizeof(crange));
    for (int v = 0; v < n_nodes && k > 0; ++v) {
        if (t.first_child[v] == -1) continue;                       /* leaves skipped :225 */
        int nl = 0;
        if (v != 0) { links[nl].start = t.hi[v] % k; links[nl].end = t.lo[v] % k; nl++; }   /* link towards the root */
        for (int c = t.first_child[v]; c != -1; c = t.next_sib[c]) { links[nl].start = t.lo[c] % k; links[nl].end = t.hi[c] % k; nl++; }
        for (int i = 0; i < nl; ++i) for (int j = i + 1; j < nl; ++j) for (int l = j + 1; l < nl; ++l) {
            three_clades(T, links[i], links[j], links[l], el, k);
            three_clades(T, links[j], links[i], links[l], el, k);
            three_clades(T, links[l], links[i], links[j], el, k);
        }
    }
    free(links); free(el); otree_free(&t);
    return 0;
}

int qso_count_clades_compact(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                             const int32_t* leaf_id, uint32_t* table) {
    ctable T = {n_taxa, (uint64_t)n_taxa * n_taxa, (uint64_t)n_taxa * n_taxa * n_taxa, NULL, table};
    for (int t = 0; t < n_trees; ++t) {
        int r = count_tree_clades(&T, (int)(node_off[t + 1] - node_off[t]), parent + node_off[t], leaf_id + node_off[t]);
        if (r) return r;
    }
    return 0;
}

int qso_count_clades_fast(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                          const int32_t* leaf_id, uint32_t* table) {
    if (n_taxa > 64) return -3;
    uint64_t n = (uint64_t)n_taxa;
    ctable T = {n_taxa, n * n, n * n * n, (uint32_t*)calloc(n * n * n * n, sizeof(uint32_t)), NULL};
    if (!T.fast) return -4;
    for (int t = 0; t < n_trees; ++t) {
        int r = count_tree_clades(&T, (int)(node_off[t + 1] - node_off[t]), parent + node_off[t], leaf_id + node_off[t]);
        if (r) { free(T.fast); return r; }
    }
#define CO(a, b, c, d) ((a) * T.n3 + (b) * T.n2 + (c) * n + (d))
#define LOOKUP(a, b, c, d) (T.fast[CO(a, b, c, d)] + T.fast[CO(a, b, d, c)] + T.fast[CO(b, a, c, d)] + T.fast[CO(b, a, d, c)]) /* :283-290 */
    for (uint64_t d = 3; d < n; ++d) for (uint64_t c = 2; c < d; ++c) for (uint64_t b = 1; b < c; ++b) for (uint64_t a = 0; a < b; ++a) {
        uint64_t r = rank_sorted(a, b, c, d);
        table[r * 3 + 0] = LOOKUP(a, b, c, d);   /* :313-315 with (a,b,c,d) sorted: ab|cd, ac|bd, ad|bc */
        table[r * 3 + 1] = LOOKUP(a, c, b, d);
        table[r * 3 + 2] = LOOKUP(a, d, b, c);
    }
    free(T.fast);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* counting: four-point restatement                                                            */
/* ------------------------------------------------------------------------------------------ */

int qso_count_fourpoint(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent,
                        const int32_t* leaf_id, uint32_t* table) {
    size_t n = (size_t)n_taxa;
    uint16_t* D = (uint16_t*)malloc(n * n * sizeof(uint16_t));
    for (int t = 0; t < n_trees; ++t) {
        int r = qso_distance_matrix((int)(node_off[t + 1] - node_off[t]), parent + node_off[t], leaf_id + node_off[t], n_taxa, D);
        if (r < 0) { free(D); return r; }
        uint64_t rank = 0;
        for (size_t d = 3; d < n; ++d) for (size_t c = 2; c < d; ++c) for (size_t b = 1; b < c; ++b) for (size_t a = 0; a < b; ++a, ++rank) {
            unsigned ab = D[a * n + b], cd = D[c * n + d], ac = D[a * n + c], bd = D[b * n + d], ad = D[a * n + d], bc = D[b * n + c];
            if (ab == QSO_MISSING || cd == QSO_MISSING || ac == QSO_MISSING || bd == QSO_MISSING || ad == QSO_MISSING || bc == QSO_MISSING) continue;
            unsigned s0 = ab + cd, s1 = ac + bd, s2 = ad + bc;
            if (s0 < s1 && s0 < s2) table[rank * 3 + 0]++;
            else if (s1 < s0 && s1 < s2) table[rank * 3 + 1]++;
            else if (s2 < s0 && s2 < s1) table[rank * 3 + 2]++;
        }
    }
    free(D);
    return 0;
}

/* Sampled restatement for sizes where the whole table is out of reach (n = 500 .. 2000): the entries of the given
 * ranks only.  A rank is unranked by inverting quartet_lookup_table.hpp:141-168 (largest s3 with C(s3,4) <= r, then
 * s2, s1, s0); per tree the six leaf-to-leaf distances are TreeInformation.hpp:40-43 evaluated by plain climbing, and the
 * topology is the unique minimum of the three pair sums.  out[i*3 + slot], caller zeroes it. */
void qso_unrank(uint64_t r, uint64_t q[4]) {
    uint64_t s3 = 3;
    while ((s3 + 1) * s3 * (s3 - 1) * (s3 - 2) / 24 <= r) ++s3;
    r -= s3 * (s3 - 1) * (s3 - 2) * (s3 - 3) / 24;
    uint64_t s2 = 2;
    while ((s2 + 1) * s2 * (s2 - 1) / 6 <= r) ++s2;
    r -= s2 * (s2 - 1) * (s2 - 2) / 6;
    uint64_t s1 = 1;
    while ((s1 + 1) * s1 / 2 <= r) ++s1;
    r -= s1 * (s1 - 1) / 2;
    q[0] = r; q[1] = s1; q[2] = s2; q[3] = s3;
}

int qso_count_fourpoint_ranks(int n_taxa, int n_trees, const int64_t* node_off, const int32_t* parent, const int32_t* leaf_id,
                              int64_t n_ranks, const uint64_t* ranks, uint32_t* out) {
    int* quads = (int*)malloc((size_t)n_ranks * 4 * sizeof(int));
    int* leaf_of = (int*)malloc((size_t)n_taxa * sizeof(int));
    for (int64_t i = 0; i < n_ranks; ++i) {
        uint64_t q[4];
        qso_unrank(ranks[i], q);
        if (q[3] >= (uint64_t)n_taxa || !(q[0] < q[1] && q[1] < q[2] && q[2] < q[3]) || rank_sorted(q[0], q[1], q[2], q[3]) != ranks[i]) { free(quads); free(leaf_of); return -5; }
        for (int k = 0; k < 4; ++k) quads[i * 4 + k] = (int)q[k];
    }
    int rc = 0;
    for (int t = 0; t < n_trees && rc == 0; ++t) {
        const int N = (int)(node_off[t + 1] - node_off[t]);
        const int32_t* lid = leaf_id + node_off[t];
        otree T;
        if (otree_build(&T, N, parent + node_off[t], lid, NULL) != 0) { otree_free(&T); rc = -1; break; }
        for (int i = 0; i < n_taxa; ++i) leaf_of[i] = -1;
        for (int i = 0; i < T.k; ++i) {
            const int v = T.leaf_seq[i];
            if (lid[v] < 0 || lid[v] >= n_taxa) { rc = -2; break; }
            leaf_of[lid[v]] = v;
        }
        if (rc == 0) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n_ranks; ++i) {
                const int a = leaf_of[quads[i * 4]], b = leaf_of[quads[i * 4 + 1]], c = leaf_of[quads[i * 4 + 2]], d = leaf_of[quads[i * 4 + 3]];
                if (a < 0 || b < 0 || c < 0 || d < 0) continue;
                const unsigned s0 = dist_edges(&T, a, b) + dist_edges(&T, c, d), s1 = dist_edges(&T, a, c) + dist_edges(&T, b, d),
                               s2 = dist_edges(&T, a, d) + dist_edges(&T, b, c);
                if (s0 < s1 && s0 < s2) out[i * 3 + 0]++;
                else if (s1 < s0 && s1 < s2) out[i * 3 + 1]++;
                else if (s2 < s0 && s2 < s1) out[i * 3 + 2]++;
            }
        }
        otree_free(&T);
    }
    free(quads); free(leaf_of);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* scoring                                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* QuartetScoreComputer.hpp:135-159, same order of operations */
double qso_log_score(uint64_t q1, uint64_t q2, uint64_t q3) {
    if (q1 == 0 && q2 == 0 && q3 == 0) return 0;
    uint64_t sum = q1 + q2 + q3;
    double p1 = (double)q1 / sum, p2 = (double)q2 / sum, p3 = (double)q3 / sum;
    double qic = 1;
    if (p1 != 0) qic += p1 * log(p1) / log(3);
    if (p2 != 0) qic += p2 * log(p2) / log(3);
    if (p3 != 0) qic += p3 * log(p3) / log(3);
    if (q1 < q2 || q1 < q3) return qic * -1;
    return qic;
}

typedef struct {
    const uint32_t* table; int scale; uint64_t mask;
} cnt_src;

/* QuartetCounterLookup.hpp:300-318 on the canonical table (lookup ids a,b,c,d, any order):
 * counts of ab|cd, ac|bd, ad|bc as the CINT the reference would have stored. */
static void count_occ(const cnt_src* S, uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t q[3]) {
    const uint32_t* tup = S->table + qso_rank(a, b, c, d) * 3;
    q[0] = ((uint64_t)tup[qso_tuple_index(a, b, c, d)] * S->scale) & S->mask;
    q[1] = ((uint64_t)tup[qso_tuple_index(a, c, b, d)] * S->scale) & S->mask;
    q[2] = ((uint64_t)tup[qso_tuple_index(a, d, b, c)] * S->scale) & S->mask;
}

/* LQ-IC update along the central path of quartet ab|cd: QuartetScoreComputer.hpp:436-454 (and :571-588) */
static void lqic_update(const otree* t, const int32_t* parent_edge, int a, int b, int c, int d, double qic, double* lqic) {
    int lca_ab = lca0(t, a, b), lca_cd = lca0(t, c, d), from, to;
    if (lca_cd == lca_r(t, c, d, lca_ab)) { from = lca_r(t, a, b, lca_cd); to = lca_cd; }
    else { from = lca_ab; to = lca_r(t, c, d, lca_ab); }
    int l = lca0(t, from, to);
    for (int x = from; x != l; x = t->parent[x]) if (qic < lqic[parent_edge[x]]) lqic[parent_edge[x]] = qic;
    for (int x = to; x != l; x = t->parent[x]) if (qic < lqic[parent_edge[x]]) lqic[parent_edge[x]] = qic;
}

/* the links of node v in genesis order: [link to parent (non-root only), children in Newick order];
 * each link identified by the cyclic leaf range behind it.  Returns count. */
static int node_links(const otree* t, int v, crange* out, int* child_of_link) {
    int nl = 0, k = t->k;
    if (v != 0) { out[nl].start = t->hi[v] % k; out[nl].end = t->lo[v] % k; child_of_link[nl] = -1; nl++; }
    for (int c = t->first_child[v]; c != -1; c = t->next_sib[c]) { out[nl].start = t->lo[c] % k; out[nl].end = t->hi[c] % k; child_of_link[nl] = c; nl++; }
    return nl;
}

/* index (in node_links order) of the link of u that points towards v: QuartetScoreComputer.hpp:169-194 */
static int link_towards(const otree* t, int u, int v, const int* child_of_link, int nl) {
    /* is v below u? climb v to depth(u)+1 */
    int x = v;
    while (t->depth[x] > t->depth[u] + 1) x = t->parent[x];
    if (t->depth[x] == t->depth[u] + 1 && t->parent[x] == u) {
        for (int i = 0; i < nl; ++i) if (child_of_link[i] == x) return i;
    }
    return 0; /* the link to the parent (primary link) */
}

/* QuartetScoreComputer.hpp:379-490 */
static void process_node_pair(const otree* t, const int32_t* parent_edge, const cnt_src* S, int u, int v,
                              double* lqic, double* qpic_out, double* eqpic) {
    crange lu[3], lv[3]; int cu[3], cv[3];
    node_links(t, u, lu, cu); node_links(t, v, lv, cv);
    int iu = link_towards(t, u, v, cu, 3), iv = link_towards(t, v, u, cv, 3);
    crange s1 = lu[(iu + 1) % 3], s2 = lu[(iu + 2) % 3], s3 = lv[(iv + 1) % 3], s4 = lv[(iv + 2) % 3];   /* :393-396 */
    int k = t->k;
    uint32_t p1 = 0, p2 = 0, p3 = 0;                                                                      /* :382 unsigned */
    for (int ai = s1.start; ai != s1.end; ai = (ai + 1) % k)
        for (int bi = s2.start; bi != s2.end; bi = (bi + 1) % k)
            for (int ci = s3.start; ci != s3.end; ci = (ci + 1) % k)
                for (int di = s4.start; di != s4.end; di = (di + 1) % k) {
                    int a = t->leaf_seq[ai], b = t->leaf_seq[bi], c = t->leaf_seq[ci], d = t->leaf_seq[di];
                    uint64_t q[3];
                    count_occ(S, (uint64_t)t->leaf_id[a], (uint64_t)t->leaf_id[b], (uint64_t)t->leaf_id[c], (uint64_t)t->leaf_id[d], q);
                    p1 += (uint32_t)q[0]; p2 += (uint32_t)q[1]; p3 += (uint32_t)q[2];
                    lqic_update(t, parent_edge, a, b, c, d, qso_log_score(q[0], q[1], q[2]), lqic);
                }
    double qpic = qso_log_score(p1, p2, p3);                                                              /* :472 */
    /* :475-481: primary link of u (for the root: its first child link) */
    int u_outer = (u == 0) ? t->first_child[0] : t->parent[u];
    int v_outer = (v == 0) ? t->first_child[0] : t->parent[v];
    if (u_outer == v) qpic_out[(u == 0) ? parent_edge[v] : parent_edge[u]] = qpic;
    else if (v_outer == u) qpic_out[(v == 0) ? parent_edge[u] : parent_edge[v]] = qpic;
    int l = lca0(t, u, v);                                                                                /* :484-489 */
    for (int x = u; x != l; x = t->parent[x]) if (qpic < eqpic[parent_edge[x]]) eqpic[parent_edge[x]] = qpic;
    for (int x = v; x != l; x = t->parent[x]) if (qpic < eqpic[parent_edge[x]]) eqpic[parent_edge[x]] = qpic;
}

/* reference topology of the leaves at Euler positions u<v<w<z: QuartetScoreComputer.hpp:526-562.
 * returns 0 (uv|wz), 1 (uw|vz), 2 (uz|vw) or -1 (unresolved) */
static int ref_topology(const otree* t, int u, int v, int w, int z) {
    int luv = lca0(t, u, v), luw = lca0(t, u, w), luz = lca0(t, u, z), lvw = lca0(t, v, w), lvz = lca0(t, v, z), lwz = lca0(t, w, z);
    unsigned d0 = dist_edges(t, luv, lwz), d1 = dist_edges(t, luw, lvz), d2 = dist_edges(t, luz, lvw);
    if (d0 > d1 && d0 > d2) return 0;
    if (d1 > d0 && d1 > d2) return 1;
    if (d2 > d0 && d2 > d1) return 2;
    return -1;
}

static uint64_t cint_mask(int bits) { return bits >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << bits) - 1); }

int qso_score(int n_nodes, const int32_t* parent, const int32_t* leaf_id, const int32_t* parent_edge,
              const int32_t* child_rank, int n_taxa, const uint32_t* table, int count_scale, int cint_bits,
              double* lqic, double* qpic, double* eqpic, int* bifurcating) {
    otree t;
    if (otree_build(&t, n_nodes, parent, leaf_id, child_rank) != 0) { otree_free(&t); return -1; }
    if (t.k != n_taxa) { otree_free(&t); return -2; }
    int E = n_nodes - 1;
    for (int e = 0; e < E; ++e) { lqic[e] = INFINITY; qpic[e] = INFINITY; eqpic[e] = INFINITY; }
    cnt_src S = {table, count_scale, cint_mask(cint_bits)};
    /* genesis functions.cpp:57-69: bifurcating <=> max rank (links - 1) == 2 */
    int max_rank = 0, bad = 0;
    for (int v = 0; v < n_nodes; ++v) {
        int links = t.n_child[v] + (v != 0 ? 1 : 0);
        if (links - 1 > max_rank) max_rank = links - 1;
        if (t.n_child[v] > 0 && links != 3) bad = 1;
    }
    *bifurcating = (max_rank == 2);
    if (*bifurcating) {
        if (bad) { otree_free(&t); return -3; }  /* the reference mis-scores these; out of contract */
        /* QuartetScoreComputer.hpp:495-508 */
        for (int i = 0; i < n_nodes; ++i) {
            if (t.n_child[i] == 0) continue;
            for (int j = i + 1; j < n_nodes; ++j) {
                if (t.n_child[j] == 0) continue;
                process_node_pair(&t, parent_edge, &S, i, j, lqic, qpic, eqpic);
            }
        }
    } else {
        /* QuartetScoreComputer.hpp:513-593 */
        for (int ui = 0; ui < t.k; ++ui) for (int vi = ui + 1; vi < t.k; ++vi) for (int wi = vi + 1; wi < t.k; ++wi) for (int zi = wi + 1; zi < t.k; ++zi) {
            int u = t.leaf_seq[ui], v = t.leaf_seq[vi], w = t.leaf_seq[wi], z = t.leaf_seq[zi];
            int topo = ref_topology(&t, u, v, w, z);
            int a, b, c, d;
            if (topo == 0) { a = u; b = v; c = w; d = z; }
            else if (topo == 1) { a = u; b = w; c = v; d = z; }
            else if (topo == 2) { a = u; b = z; c = v; d = w; }
            else continue;
            uint64_t q[3];
            count_occ(&S, (uint64_t)leaf_id[a], (uint64_t)leaf_id[b], (uint64_t)leaf_id[c], (uint64_t)leaf_id[d], q);
            lqic_update(&t, parent_edge, a, b, c, d, qso_log_score(q[0], q[1], q[2]), lqic);
        }
    }
    otree_free(&t);
    return 0;
}

int qso_raw_qic(int n_nodes, const int32_t* parent, const int32_t* leaf_id, int n_taxa,
                const uint32_t* table, int count_scale, int cint_bits, int8_t* topo_out, double* qic_out) {
    otree t;
    if (otree_build(&t, n_nodes, parent, leaf_id, NULL) != 0) { otree_free(&t); return -1; }
    /* NOTE: leaf ids are positions in Euler order by construction, so sibling order by index is only
       valid if the caller numbered nodes accordingly; use the leaf ids to locate leaves instead. */
    int* leaf_of = (int*)malloc(n_taxa * sizeof(int));
    for (int i = 0; i < n_taxa; ++i) leaf_of[i] = -1;
    for (int i = 0; i < n_nodes; ++i) if (t.first_child[i] == -1 && leaf_id[i] >= 0 && leaf_id[i] < n_taxa) leaf_of[leaf_id[i]] = i;
    for (int i = 0; i < n_taxa; ++i) if (leaf_of[i] < 0) { free(leaf_of); otree_free(&t); return -2; }
    cnt_src S = {table, count_scale, cint_mask(cint_bits)};
    size_t idx = 0;
    for (int ui = 0; ui < n_taxa; ++ui) for (int vi = ui + 1; vi < n_taxa; ++vi) for (int wi = vi + 1; wi < n_taxa; ++wi) for (int zi = wi + 1; zi < n_taxa; ++zi, ++idx) {
        int u = leaf_of[ui], v = leaf_of[vi], w = leaf_of[wi], z = leaf_of[zi];
        int topo = ref_topology(&t, u, v, w, z);
        topo_out[idx] = (int8_t)topo; qic_out[idx] = 0;
        int a, b, c, d;
        if (topo == 0) { a = ui; b = vi; c = wi; d = zi; }
        else if (topo == 1) { a = ui; b = wi; c = vi; d = zi; }
        else if (topo == 2) { a = ui; b = zi; c = vi; d = wi; }
        else continue;
        uint64_t q[3];
        count_occ(&S, (uint64_t)a, (uint64_t)b, (uint64_t)c, (uint64_t)d, q);
        qic_out[idx] = qso_log_score(q[0], q[1], q[2]);
    }
    free(leaf_of); otree_free(&t);
    return 0;
}
