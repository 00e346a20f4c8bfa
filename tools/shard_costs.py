#!/usr/bin/env python3
"""Counting-kernel time of every shard of a G-way split, one after the other on ONE GPU (same clocks, same everything):
ground truth for the shard cost model in qscuda.cu shard_bounds."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500); ap.add_argument("--m", type=int, default=1000); ap.add_argument("--seed", type=int, default=3000)
ap.add_argument("--p-missing", type=float, default=0.1); ap.add_argument("--p-contract", type=float, default=0.05); ap.add_argument("--G", type=int, default=8)
ap.add_argument("--rebalance", type=int, default=1)
args = ap.parse_args()
from quartetscores_b200 import Context, QS_MODE_AUTO
from quartetscores_b200.computer import cint_bytes_for
from quartetscores_b200.newick import flatten_reference, parse_newick
from quartetscores_b200.synth import SyntheticInput
s = SyntheticInput(args.n, args.m, args.seed, k_max=20, p_missing=args.p_missing, p_contract=args.p_contract, want_newick=False)
ref = flatten_reference(parse_newick(s.ref_newick))
for g in range(args.G):
    with Context(args.n, cint_bytes_for(args.m), mode=QS_MODE_AUTO, shard_index=g, shard_count=args.G) as ctx:
        ctx.set_reference(ref); ctx.add_trees(s.flat)
        if args.rebalance: ctx.rebalance_shards()
        best = 1e9
        for _ in range(3):
            ctx.count(); best = min(best, ctx.last_timing()["count_ms"])
        r0, r1 = ctx.shard_range()
        ctx.score(1)
        print(f"n={args.n} m={args.m} G={args.G} shard={g} quartets={r1 - r0} count_ms={best:.3f} score_ms={ctx.last_timing()['score_ms']:.3f}", flush=True)
