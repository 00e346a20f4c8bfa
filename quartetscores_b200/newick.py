"""Newick reading/writing and tree flattening for the host side of the C ABI.

Host-side mirror of what the reference gets from genesis (reference files cited relative to
/root/reference):

* parsing: genesis/lib/genesis/tree/formats/newick/reader.cpp:397-812 (names, ``:length``,
  ``[comments]``, quoted labels, several trees per file separated by ``;``);
* node numbering of the REFERENCE tree: root = 0, then pre-order with the children of every node
  visited in REVERSE Newick order, the edge above node i has index i-1 (reader.cpp:591-758,
  SURVEY.md Appendix A1) -- per-edge outputs are indexed by that edge index;
* taxon (lookup) ids: position of the leaf in the reference tree's Euler-tour leaf order = left to
  right in the Newick text (src/QuartetCounterLookup.hpp:249-258);
* writing: genesis/lib/genesis/tree/formats/newick/writer.cpp:58-198 plus the quartet plugin
  src/quartet_newick_writer.hpp:160-188 (``[qp-ic:..;lq-ic:..;eqp-ic:..]`` comments).

Everything here is plain Python/numpy host plumbing; no GPU work.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np


@dataclass
class Node:
    name: str = ""
    length: Optional[str] = None          # branch length text as read (None if absent)
    children: List["Node"] = field(default_factory=list)

    def is_leaf(self) -> bool:
        return not self.children


class NewickError(ValueError):
    pass


def _tokenize(text: str) -> Iterator[tuple]:
    i, n = 0, len(text)
    while i < n:
        ch = text[i]
        if ch.isspace():
            i += 1
        elif ch in "(),:;":
            yield (ch, ch)
            i += 1
        elif ch == "[":
            j = text.find("]", i)
            if j < 0:
                raise NewickError("unterminated comment")
            i = j + 1                       # comments are dropped (the reference ignores them)
        elif ch in "'\"":                   # genesis reader.cpp:371-373: either quote, a doubled quote stands for itself
            j = i + 1
            out = []
            while True:
                k = text.find(ch, j)
                if k < 0:
                    raise NewickError("unterminated quoted label")
                out.append(text[j:k])
                if k + 1 < n and text[k + 1] == ch:
                    out.append(ch)
                    j = k + 2
                else:
                    i = k + 1
                    break
            yield ("label", "".join(out))
        else:
            j = i
            while j < n and not text[j].isspace() and text[j] not in "(),:;[]":
                j += 1
            yield ("label", text[i:j])
            i = j


def parse_newick_many(text: str) -> List[Node]:
    """Parse every tree in `text` (iteratively, so deep caterpillars do not hit the recursion limit)."""
    trees: List[Node] = []
    stack: List[Node] = []
    cur: Optional[Node] = None            # node whose label/length may still follow
    expect_length = False
    started = False
    for kind, val in _tokenize(text):
        if kind == "(":
            node = Node()
            if stack:
                stack[-1].children.append(node)
            stack.append(node)
            cur = None
            started = True
        elif kind == ",":
            if not stack:
                raise NewickError("',' outside of a tree")
            if cur is None:                # empty leaf name
                stack[-1].children.append(Node())
            cur = None
            # a following label starts a new leaf of stack[-1]
        elif kind == ")":
            if not stack:
                raise NewickError("unbalanced ')'")
            if cur is None:
                stack[-1].children.append(Node())
            cur = stack.pop()
            cur._closed = True              # type: ignore[attr-defined]
        elif kind == ":":
            if cur is None:
                cur = Node()
                if stack:
                    stack[-1].children.append(cur)
            expect_length = True
        elif kind == "label":
            if expect_length:
                cur.length = val            # type: ignore[union-attr]
                expect_length = False
            elif cur is not None and getattr(cur, "_closed", False) and not cur.name:
                cur.name = val              # inner node label
            else:
                cur = Node(name=val)
                if stack:
                    stack[-1].children.append(cur)
                started = True
        elif kind == ";":
            if stack:
                raise NewickError("unbalanced '(' at ';'")
            if cur is None:
                if started:
                    raise NewickError("empty tree")
                continue
            trees.append(cur)
            cur, started, expect_length = None, False, False
    if stack or (cur is not None and started):
        raise NewickError("tree not terminated by ';'")
    return trees


def parse_newick(text: str) -> Node:
    t = parse_newick_many(text)
    if len(t) != 1:
        raise NewickError(f"expected one tree, found {len(t)}")
    return t[0]


# ---------------------------------------------------------------------------------------------
# flattening
# ---------------------------------------------------------------------------------------------

@dataclass
class FlatReference:
    """Reference tree in the C-ABI encoding (include/qscuda.h: qs_set_reference)."""
    parent: np.ndarray          # int32[n_nodes], -1 for the root
    parent_edge: np.ndarray     # int32[n_nodes], genesis edge index above the node, -1 for the root
    leaf_lookup_id: np.ndarray  # int32[n_nodes], -1 for inner nodes
    first_child: np.ndarray     # int32[n_nodes]
    next_sibling: np.ndarray    # int32[n_nodes]
    child_rank: np.ndarray      # int32[n_nodes] position among siblings in Newick order
    names: List[str]            # per node
    lengths: List[Optional[str]]
    taxa: List[str]             # taxon name per lookup id

    @property
    def n_nodes(self) -> int:
        return len(self.parent)

    @property
    def n_taxa(self) -> int:
        return len(self.taxa)

    @property
    def edge_count(self) -> int:
        return len(self.parent) - 1


def flatten_reference(root: Node) -> FlatReference:
    nodes: List[Node] = []
    parent: List[int] = []
    rank: List[int] = []
    # genesis numbering: pre-order, children in reverse Newick order
    stack = [(root, -1, 0)]
    while stack:
        node, par, rk = stack.pop()
        idx = len(nodes)
        nodes.append(node)
        parent.append(par)
        rank.append(rk)
        # push in Newick order so that the LAST child is popped (numbered) first
        for r, ch in enumerate(node.children):
            stack.append((ch, idx, r))
    n = len(nodes)
    index_of = {id(nd): i for i, nd in enumerate(nodes)}
    first_child = np.full(n, -1, np.int32)
    next_sibling = np.full(n, -1, np.int32)
    for i, nd in enumerate(nodes):
        prev = -1
        for ch in nd.children:
            ci = index_of[id(ch)]
            if prev < 0:
                first_child[i] = ci
            else:
                next_sibling[prev] = ci
            prev = ci
    # lookup ids: leaves left to right
    leaf_lookup = np.full(n, -1, np.int32)
    taxa: List[str] = []
    st = [root]
    while st:
        nd = st.pop()
        if nd.is_leaf():
            leaf_lookup[index_of[id(nd)]] = len(taxa)
            taxa.append(nd.name)
        else:
            st.extend(reversed(nd.children))
    if len(set(taxa)) != len(taxa):
        raise NewickError("duplicate taxon names in the reference tree")
    par = np.asarray(parent, np.int32)
    return FlatReference(
        parent=par,
        parent_edge=np.where(par >= 0, np.arange(n, dtype=np.int32) - 1, -1).astype(np.int32),
        leaf_lookup_id=leaf_lookup,
        first_child=first_child,
        next_sibling=next_sibling,
        child_rank=np.asarray(rank, np.int32),
        names=[nd.name for nd in nodes],
        lengths=[nd.length for nd in nodes],
        taxa=taxa,
    )


@dataclass
class FlatTrees:
    """Evaluation (gene) trees in the C-ABI encoding (include/qscuda.h: qs_add_trees)."""
    node_offsets: np.ndarray    # int64[n_trees + 1]
    parent: np.ndarray          # int32[total nodes], parent[i] < i inside a tree, -1 for a root
    leaf_lookup_id: np.ndarray  # int32[total nodes], -1 for inner nodes

    @property
    def n_trees(self) -> int:
        return len(self.node_offsets) - 1

    def slice(self, begin: int, end: int) -> "FlatTrees":
        o = self.node_offsets
        return FlatTrees(o[begin:end + 1] - o[begin], self.parent[o[begin]:o[end]], self.leaf_lookup_id[o[begin]:o[end]])


def flatten_eval_tree(root: Node, taxon_to_id: Dict[str, int]):
    """Pre-order flattening of one evaluation tree.  Unknown taxa raise KeyError exactly where the
    reference throws (src/QuartetCounterLookup.hpp:218)."""
    parent: List[int] = []
    leaf: List[int] = []
    stack = [(root, -1)]
    while stack:
        node, par = stack.pop()
        idx = len(parent)
        parent.append(par)
        if node.is_leaf():
            leaf.append(taxon_to_id[node.name])
        else:
            leaf.append(-1)
            for ch in reversed(node.children):
                stack.append((ch, idx))
    return parent, leaf


def flatten_eval_trees(roots: Sequence[Node], taxa: Sequence[str]) -> FlatTrees:
    t2i = {name: i for i, name in enumerate(taxa)}
    offs = [0]
    par_all: List[int] = []
    leaf_all: List[int] = []
    for r in roots:
        p, l = flatten_eval_tree(r, t2i)
        par_all.extend(p)
        leaf_all.extend(l)
        offs.append(len(par_all))
    return FlatTrees(np.asarray(offs, np.int64), np.asarray(par_all, np.int32), np.asarray(leaf_all, np.int32))


# ---------------------------------------------------------------------------------------------
# writing
# ---------------------------------------------------------------------------------------------

def _to_string_rounded(value: float, precision: int = 6) -> str:
    """genesis utils/text/string.cpp:398-414: fixed notation, trailing zeros (and a trailing '.') removed."""
    s = f"{value:.{precision}f}"
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


def _quote(name: str) -> str:
    # writer.cpp:132-141
    if any(ch in name for ch in " :;()[],"):
        return '"' + name + '"'
    return name


def write_annotated_newick(root: Node, flat: FlatReference, lqic, qpic=None, eqpic=None) -> str:
    """Annotated Newick exactly as QuartetTreeNewickWriter produces it (no trailing newline).

    Branch lengths are rewritten as genesis does (default 1.0 when absent, rounded to 6 decimals,
    genesis/lib/genesis/tree/default/newick_reader.hpp:297, newick_writer.hpp:287-294); the comment
    carries ``qp-ic`` / ``lq-ic`` / ``eqp-ic`` in that order, each only where lq-ic is finite
    (quartet_newick_writer.hpp:164-187; the reference tests lq-ic for all three).
    """
    index_of: Dict[int, int] = {}
    # recompute node -> index with the same traversal as flatten_reference
    stack = [root]
    order: List[Node] = []
    while stack:
        nd = stack.pop()
        index_of[id(nd)] = len(order)
        order.append(nd)
        for ch in nd.children:
            stack.append(ch)

    def element(nd: Node, is_root: bool) -> str:
        res = _quote(nd.name)
        if is_root:
            return res
        length = 1.0 if nd.length is None else float(nd.length)
        res += ":" + _to_string_rounded(length)
        e = index_of[id(nd)] - 1
        parts = []
        finite = lqic[e] != float("inf")
        if qpic is not None and finite:
            parts.append("qp-ic:%f" % qpic[e])
        if finite:
            parts.append("lq-ic:%f" % lqic[e])
        if eqpic is not None and finite:
            parts.append("eqp-ic:%f" % eqpic[e])
        if parts:
            res += "[" + ";".join(parts) + "]"
        return res

    # iterative post-order string assembly
    out: Dict[int, str] = {}
    st = [(root, False)]
    while st:
        nd, done = st.pop()
        if nd.is_leaf():
            out[id(nd)] = element(nd, nd is root)
        elif done:
            out[id(nd)] = "(" + ",".join(out.pop(id(ch)) for ch in nd.children) + ")" + element(nd, nd is root)
        else:
            st.append((nd, True))
            for ch in nd.children:
                st.append((ch, False))
    return out[id(root)] + ";"


def to_newick(root: Node) -> str:
    """Plain Newick (names and lengths as stored), for writing synthetic inputs."""
    out: Dict[int, str] = {}
    st = [(root, False)]
    while st:
        nd, done = st.pop()
        suffix = nd.name + ("" if nd.length is None else ":" + nd.length)
        if nd.is_leaf():
            out[id(nd)] = suffix
        elif done:
            out[id(nd)] = "(" + ",".join(out.pop(id(ch)) for ch in nd.children) + ")" + suffix
        else:
            st.append((nd, True))
            for ch in nd.children:
                st.append((ch, False))
    return out[id(root)] + ";"
