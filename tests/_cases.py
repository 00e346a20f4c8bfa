"""Shared helpers: turn a golden fixture (Newick text) into the flat C-ABI encoding."""
from quartetscores_b200.newick import flatten_eval_trees, flatten_reference, parse_newick, parse_newick_many


def load_input(g):
    ref_root = parse_newick(g["ref_newick"])
    ref = flatten_reference(ref_root)
    evals = parse_newick_many(g["eval_newick"])
    flat = flatten_eval_trees(evals, ref.taxa)
    return ref_root, ref, flat


def cint_bits_for(m):
    # src/QuartetScores.cpp:115-147
    return 8 if m < (1 << 8) else 16 if m < (1 << 16) else 32 if m < (1 << 32) else 64
