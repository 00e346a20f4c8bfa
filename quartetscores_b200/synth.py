"""Seeded synthetic inputs for tests and bench.py (SURVEY.md §8d "concrete synthetic inputs").

* reference tree: random unrooted binary tree (trifurcating root) on taxa ``t0..t{n-1}``;
* gene trees: the reference topology after ``k ~ U{0..k_max}`` random SPR moves (an NNI is the
  SPR to a neighbouring edge), optionally with taxa dropped independently (missing data) and
  internal edges contracted (multifurcations).

Trees are manipulated as child lists over integer node ids and emitted either as Newick text
(what the reference binary reads) or directly in the flat C-ABI encoding (newick.FlatTrees).
Pure host-side Python; deterministic for a given seed.
"""
from __future__ import annotations

import random
from typing import List, Optional, Tuple

import numpy as np

from .newick import FlatTrees, Node


class _T:
    """Mutable rooted tree: children lists + parent pointers; leaves carry a taxon id."""

    __slots__ = ("children", "parent", "taxon", "root")

    def __init__(self):
        self.children: List[List[int]] = []
        self.parent: List[int] = []
        self.taxon: List[int] = []
        self.root = 0

    def new(self, taxon: int = -1) -> int:
        self.children.append([])
        self.parent.append(-1)
        self.taxon.append(taxon)
        return len(self.parent) - 1

    def attach(self, child: int, par: int) -> None:
        self.children[par].append(child)
        self.parent[child] = par

    def detach(self, child: int) -> None:
        p = self.parent[child]
        self.children[p].remove(child)
        self.parent[child] = -1

    def copy(self) -> "_T":
        t = _T()
        t.children = [list(c) for c in self.children]
        t.parent = list(self.parent)
        t.taxon = list(self.taxon)
        t.root = self.root
        return t

    def live_nodes(self) -> List[int]:
        out, st = [], [self.root]
        while st:
            v = st.pop()
            out.append(v)
            st.extend(self.children[v])
        return out


def random_reference(n: int, rng: random.Random) -> _T:
    """Random unrooted binary topology by stepwise addition; root of degree 3."""
    assert n >= 4
    order = list(range(n))
    rng.shuffle(order)
    t = _T()
    root = t.new()
    for x in order[:3]:
        t.attach(t.new(x), root)
    edges = list(t.children[root])                 # an edge is identified by its lower node
    for x in order[3:]:
        y = rng.choice(edges)
        p = t.parent[y]
        mid = t.new()
        idx = t.children[p].index(y)
        t.children[p][idx] = mid
        t.parent[mid] = p
        t.children[mid] = []
        t.attach(y, mid)
        leaf = t.new(x)
        if rng.random() < 0.5:
            t.attach(leaf, mid)
        else:
            t.children[mid].insert(0, leaf)
            t.parent[leaf] = mid
        edges.extend([mid, leaf])
    return t


def _in_subtree(t: _T, node: int, top: int) -> bool:
    while node != -1:
        if node == top:
            return True
        node = t.parent[node]
    return False


def spr_move(t: _T, rng: random.Random, nni: bool = False) -> None:
    """One random subtree-prune-and-regraft (or NNI) that keeps the root node in place."""
    live = t.live_nodes()
    nodes = [v for v in live if v != t.root and t.parent[v] != t.root]
    if not nodes:
        return
    for _ in range(20):
        x = rng.choice(nodes)
        p = t.parent[x]
        g = t.parent[p]
        if len(t.children[p]) != 2:
            continue
        sib = t.children[p][0] if t.children[p][1] == x else t.children[p][1]
        sub = set()                                # nodes of the pruned subtree (one DFS instead of a root walk per candidate)
        st = [x]
        while st:
            v = st.pop()
            sub.add(v)
            st.extend(t.children[v])
        if nni:
            cands = [c for c in t.children[g] if c != p] + ([g] if g != t.root else [])
        else:
            cands = [v for v in live if v != t.root and v != p and v != sib and v not in sub]
        cands = [v for v in cands if v != t.root and v not in sub and v != p]
        if not cands:
            continue
        y = rng.choice(cands)
        # prune: splice p out
        idx = t.children[g].index(p)
        t.children[g][idx] = sib
        t.parent[sib] = g
        # regraft p on the edge above y
        py = t.parent[y]
        idy = t.children[py].index(y)
        t.children[py][idy] = p
        t.parent[p] = py
        t.children[p] = [y, x] if rng.random() < 0.5 else [x, y]
        t.parent[y] = p
        t.parent[x] = p
        return


def drop_taxa(t: _T, p_missing: float, rng: random.Random, keep_min: int = 0) -> None:
    leaves = [v for v in t.live_nodes() if not t.children[v]]
    drop = [v for v in leaves if rng.random() < p_missing]
    if len(leaves) - len(drop) < keep_min:
        drop = drop[: max(0, len(leaves) - keep_min)]
    for v in drop:
        p = t.parent[v]
        t.detach(v)
        # suppress unary / empty inner nodes upwards
        while p != -1 and p != t.root and len(t.children[p]) <= 1:
            g = t.parent[p]
            if len(t.children[p]) == 1:
                c = t.children[p][0]
                idx = t.children[g].index(p)
                t.children[g][idx] = c
                t.parent[c] = g
            else:
                t.children[g].remove(p)
            p = g
    while len(t.children[t.root]) == 1 and t.children[t.children[t.root][0]]:
        t.root = t.children[t.root][0]
        t.parent[t.root] = -1


def contract_edges(t: _T, p_contract: float, rng: random.Random) -> None:
    for v in t.live_nodes():
        if v == t.root or not t.children[v] or t.parent[v] == -1:
            continue
        if rng.random() < p_contract:
            p = t.parent[v]
            idx = t.children[p].index(v)
            t.children[p][idx:idx + 1] = t.children[v]
            for c in t.children[v]:
                t.parent[c] = p
            t.children[v] = []
            t.parent[v] = -1


def to_node(t: _T, names: List[str]) -> Node:
    made = {}
    order = t.live_nodes()
    for v in reversed(order):
        made[v] = Node(name=names[t.taxon[v]] if t.taxon[v] >= 0 else "", children=[made[c] for c in t.children[v]])
    return made[t.root]


def newick_of(t: _T, names: List[str]) -> str:
    out = {}
    for v in reversed(t.live_nodes()):
        if t.children[v]:
            out[v] = "(" + ",".join(out.pop(c) for c in t.children[v]) + ")"
        else:
            out[v] = names[t.taxon[v]]
    return out[t.root] + ";"


def flatten(t: _T, taxon_to_lookup: List[int]) -> Tuple[List[int], List[int]]:
    """Pre-order flat encoding (parent[i] < i)."""
    parent, leaf = [], []
    st = [(t.root, -1)]
    while st:
        v, par = st.pop()
        idx = len(parent)
        parent.append(par)
        if t.children[v]:
            leaf.append(-1)
            for c in reversed(t.children[v]):
                st.append((c, idx))
        else:
            leaf.append(taxon_to_lookup[t.taxon[v]])
    return parent, leaf


class SyntheticInput:
    """One seeded instance: reference tree + gene trees, as Newick and as flat arrays."""

    def __init__(self, n_taxa: int, n_trees: int, seed: int, k_max: int = 10, p_missing: float = 0.0,
                 p_contract: float = 0.0, nni_fraction: float = 0.5, want_newick: bool = True,
                 multifurcating_reference: float = 0.0):
        rng = random.Random(seed)
        self.n_taxa, self.n_trees, self.seed = n_taxa, n_trees, seed
        self.names = [f"t{i}" for i in range(n_taxa)]
        ref = random_reference(n_taxa, rng)
        if multifurcating_reference > 0:
            contract_edges(ref, multifurcating_reference, rng)
        self.ref_tree = ref
        self.ref_newick = newick_of(ref, self.names)
        # lookup id = position in the reference's left-to-right leaf order
        order = [v for v in self._leaf_order(ref)]
        self.taxon_to_lookup = [0] * n_taxa
        for pos, v in enumerate(order):
            self.taxon_to_lookup[ref.taxon[v]] = pos
        self.taxa_by_lookup = [self.names[ref.taxon[v]] for v in order]
        offs, par_all, leaf_all, nwk = [0], [], [], []
        for _ in range(n_trees):
            g = ref.copy()
            for _ in range(rng.randint(0, k_max)):
                spr_move(g, rng, nni=rng.random() < nni_fraction)
            if p_missing > 0:
                drop_taxa(g, p_missing, rng, keep_min=1)
            if p_contract > 0:
                contract_edges(g, p_contract, rng)
            p, l = flatten(g, self.taxon_to_lookup)
            par_all.extend(p)
            leaf_all.extend(l)
            offs.append(len(par_all))
            if want_newick:
                nwk.append(newick_of(g, self.names))
        self.eval_newick = nwk
        self.flat = FlatTrees(np.asarray(offs, np.int64), np.asarray(par_all, np.int32), np.asarray(leaf_all, np.int32))

    @staticmethod
    def _leaf_order(t: _T) -> List[int]:
        out, st = [], [t.root]
        while st:
            v = st.pop()
            if t.children[v]:
                st.extend(reversed(t.children[v]))
            else:
                out.append(v)
        return out

    def write(self, ref_path: str, eval_path: str) -> None:
        with open(ref_path, "w") as f:
            f.write(self.ref_newick + "\n")
        with open(eval_path, "w") as f:
            f.write("\n".join(self.eval_newick) + "\n")
