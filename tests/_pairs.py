"""Test helper: the four leaf sets around an inner-node pair of a bifurcating reference tree, as
QuartetScoreComputer::processNodePair forms them (src/QuartetScoreComputer.hpp:393-410), from the flat encoding."""
import numpy as np

import _oracle as O


class RefPairs:
    def __init__(self, ref, inner_nodes=None):
        """inner_nodes: Context.inner_nodes() (the library's own order); default: the documented rule restated here"""
        self.ref = ref
        N, n = ref.n_nodes, ref.n_taxa
        lo, hi = np.full(N, 1 << 30), np.full(N, -1)
        for v in range(N - 1, -1, -1):                 # parent[i] < i: children are finished before their parent
            if ref.leaf_lookup_id[v] >= 0:
                lo[v], hi[v] = ref.leaf_lookup_id[v], ref.leaf_lookup_id[v] + 1
            p = ref.parent[v]
            if p >= 0:
                lo[p], hi[p] = min(lo[p], lo[v]), max(hi[p], hi[v])
        self.lo, self.hi = lo, hi
        self.inner = [v for v in range(N) if ref.leaf_lookup_id[v] < 0]
        self.I = len(self.inner)
        self.kids = {v: [] for v in self.inner}
        for c in range(1, N):
            self.kids[ref.parent[c]].append(c)
        for v in self.inner:
            self.kids[v].sort(key=lambda c: lo[c])                         # planar (Newick) order
        # inner index = rank of the node's first gap in the planar leaf order (qscuda.cu build_reference); single-child nodes last
        key = {v: (hi[self.kids[v][0]] - 1 if len(self.kids[v]) >= 2 else n + v) for v in self.inner}
        self.iidx = {v: i for i, v in enumerate(sorted(self.inner, key=lambda v: key[v]))}
        if inner_nodes is not None:
            assert [int(v) for v in inner_nodes] == sorted(self.inner, key=lambda v: key[v]), "inner-node order differs from the documented rule"
        self.full = frozenset(range(n))

    def link_sets(self, x):
        """leaf sets behind the links of inner node x: its children, then (non-root) everything outside its subtree"""
        out = [frozenset(range(self.lo[c], self.hi[c])) for c in self.kids[x]]
        if self.ref.parent[x] >= 0:
            out.append(self.full - frozenset(range(self.lo[x], self.hi[x])))
        return out

    def sets(self, u, v):
        """[S1, S2, S3, S4]: the two side subtrees at u and the two at v (sorted id lists), or None if a node is not of degree 3"""
        lu, lv = self.link_sets(u), self.link_sets(v)
        if len(lu) != 3 or len(lv) != 3:
            return None

        def toward(ls, other):                         # the link that contains both side subtrees of the other node
            for i, s in enumerate(ls):
                if sum(1 for o in other if o <= s) >= 2:
                    return i
            return None

        iu, iv = toward(lu, lv), toward(lv, lu)
        if iu is None or iv is None:
            return None
        return [sorted(x) for k, x in enumerate(lu) if k != iu] + [sorted(x) for k, x in enumerate(lv) if k != iv]

    @staticmethod
    def quartets(S):
        """ranks of all quartets (a in S1, b in S2, c in S3, d in S4) and, per quartet, the table slots of ab|cd, ac|bd, ad|bc"""
        ranks, slots = [], []
        for a in S[0]:
            for b in S[1]:
                for c in S[2]:
                    for d in S[3]:
                        ranks.append(O.rank(a, b, c, d))
                        slots.append((O.tuple_index(a, b, c, d), O.tuple_index(a, c, b, d), O.tuple_index(a, d, b, c)))
        return ranks, slots

    @staticmethod
    def sums(counts, slots):
        """(p1, sorted(p2, p3)) of one pair from the counts [k,3] of its quartets"""
        c, k = np.asarray(counts, np.uint64), np.asarray(slots)
        rows = np.arange(len(c))
        return int(c[rows, k[:, 0]].sum()), sorted((int(c[rows, k[:, 1]].sum()), int(c[rows, k[:, 2]].sum())))
