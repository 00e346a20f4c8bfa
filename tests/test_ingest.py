"""Native Newick ingest (csrc/ingest.cpp): qs_newick_flatten must produce exactly the arrays of the
Python host mirror (newick.flatten_eval_trees), for any thread count, and reject what the reference rejects.
CPU only: the parser needs no GPU and no context."""
import numpy as np
import pytest

from quartetscores_b200 import QSError, flatten_newick_native
from quartetscores_b200.newick import flatten_eval_trees, flatten_reference, parse_newick, parse_newick_many
from quartetscores_b200.synth import SyntheticInput


def _same(a, b):
    assert np.array_equal(a.node_offsets, b.node_offsets)
    assert np.array_equal(a.parent, b.parent)
    assert np.array_equal(a.leaf_lookup_id, b.leaf_lookup_id)


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("cfg", [dict(n_taxa=16, n_trees=300, seed=5, k_max=6), dict(n_taxa=40, n_trees=700, seed=9, k_max=12, p_missing=0.15, p_contract=0.1)])
def test_native_flatten_matches_python(cfg, threads):
    s = SyntheticInput(want_newick=True, **cfg)
    ref = flatten_reference(parse_newick(s.ref_newick))
    text = "\n".join(s.eval_newick) + "\n"
    want = flatten_eval_trees(parse_newick_many(text), ref.taxa)
    _same(flatten_newick_native(text, ref.taxa, threads), want)



def test_native_flatten_grammar():
    taxa = ["A", "B b", "C'c", "D", "E"]
    text = """
      (A:0.1,'B b':2e-3,('C''c',D)x:1.5)root;   [a comment; with a semicolon]
      ((A,D)[inner comment]lbl:3,E , "B b");
      A;
      ( A , ( D , E ) ) ;"""
    want = flatten_eval_trees(parse_newick_many(text), taxa)
    got = flatten_newick_native(text, taxa, 2)
    _same(got, want)
    assert got.n_trees == 4
    assert list(got.leaf_lookup_id[: got.node_offsets[1]]) == [-1, 0, 1, -1, 2, 3]
    assert list(got.parent[: got.node_offsets[1]]) == [-1, 0, 0, 0, 3, 3]


def test_native_flatten_empty_and_blank():
    assert flatten_newick_native("", ["A"]).n_trees == 0
    assert flatten_newick_native("  \n ; ;\n", ["A"]).n_trees == 0


@pytest.mark.parametrize("text,needle", [
    ("(A,B,Z);", "taxon 'Z'"),                 # reference: std::out_of_range (QuartetCounterLookup.hpp:218)
    ("(A,B));", "unbalanced ')'"),
    ("((A,B);", "unbalanced '('"),
    ("(A,B)", "not terminated"),
    ("(A,'B);", "unterminated quoted"),
    ("(A,B)[oops;", "unterminated comment"),
    ("(A,B)(C,D);", "'(' directly after"),
])
def test_native_flatten_errors(text, needle):
    with pytest.raises(QSError) as e:
        flatten_newick_native(text, ["A", "B", "C", "D"])
    assert needle in str(e.value)


def test_error_names_first_bad_tree_regardless_of_threads():
    good = "(A,B,(C,D));\n"
    text = good * 200 + "(A,B,(C,Q));\n" + good * 200 + "(A,B,(C,R));\n"
    for th in (1, 4):
        with pytest.raises(QSError) as e:
            flatten_newick_native(text, ["A", "B", "C", "D"], th)
        assert "evaluation tree 200" in str(e.value) and "'Q'" in str(e.value)
