#!/usr/bin/env python3
"""Run qs_count (+ score) a few times on one synthetic workload: the command ncu wraps (tools/gpu_profile.sh)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--m", type=int, default=10000)
ap.add_argument("--seed", type=int, default=2000)
ap.add_argument("--k-max", type=int, default=20)
ap.add_argument("--p-missing", type=float, default=0.0)
ap.add_argument("--p-contract", type=float, default=0.0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--table-free", action="store_true")
args = ap.parse_args()

from quartetscores_b200 import Context, QS_MODE_TABLE, QS_MODE_TABLE_FREE
from quartetscores_b200.computer import cint_bytes_for
from quartetscores_b200.newick import flatten_reference, parse_newick
from quartetscores_b200.synth import SyntheticInput

s = SyntheticInput(args.n, args.m, args.seed, k_max=args.k_max, p_missing=args.p_missing, p_contract=args.p_contract, want_newick=False)
ref = flatten_reference(parse_newick(s.ref_newick))
with Context(args.n, cint_bytes_for(args.m), mode=QS_MODE_TABLE_FREE if args.table_free else QS_MODE_TABLE) as ctx:
    ctx.set_reference(ref)
    ctx.add_trees(s.flat)
    for _ in range(args.reps):
        ctx.count()
        ctx.score(1)
        print(ctx.last_timing(), ctx.tree_classes(), flush=True)
