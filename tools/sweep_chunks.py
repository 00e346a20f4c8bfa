#!/usr/bin/env python3
"""Sweep the tree-chunk count of the counting kernel (QS_CHUNK_COUNT tuning hook) on one workload; prints count_ms."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100); ap.add_argument("--m", type=int, default=10000); ap.add_argument("--seed", type=int, default=2000)
ap.add_argument("--p-missing", type=float, default=0.0); ap.add_argument("--p-contract", type=float, default=0.0)
ap.add_argument("--chunks", default="0,3,4,5,6,7,8,10,12,16")
ap.add_argument("--ring", default="8x8,4x8,16x8,16x4,32x2,8x4,8x2,2x8", help="trees-per-stage x stages caps to try (QS_MAX_TPS / QS_MAX_STAGES)")
args = ap.parse_args()
from quartetscores_b200 import Context
from quartetscores_b200.computer import cint_bytes_for
from quartetscores_b200.newick import flatten_reference, parse_newick
from quartetscores_b200.synth import SyntheticInput
s = SyntheticInput(args.n, args.m, args.seed, k_max=20, p_missing=args.p_missing, p_contract=args.p_contract, want_newick=False)
ref = flatten_reference(parse_newick(s.ref_newick))
with Context(args.n, cint_bytes_for(args.m)) as ctx:
    ctx.set_reference(ref); ctx.add_trees(s.flat)
    for k in [int(x) for x in args.chunks.split(",")]:
        if k: os.environ["QS_CHUNK_COUNT"] = str(k)
        else: os.environ.pop("QS_CHUNK_COUNT", None)
        best = 1e9
        for _ in range(5):
            ctx.count(); best = min(best, ctx.last_timing()["count_ms"])
        print(f"n={args.n} m={args.m} chunks={k or 'auto'} count_ms={best:.3f}", flush=True)
    os.environ.pop("QS_CHUNK_COUNT", None)
    for r in args.ring.split(","):
        tps, st = r.split("x")
        os.environ["QS_MAX_TPS"], os.environ["QS_MAX_STAGES"] = tps, st
        best = 1e9
        for _ in range(5):
            ctx.count(); best = min(best, ctx.last_timing()["count_ms"])
        print(f"n={args.n} m={args.m} ring={r} count_ms={best:.3f}", flush=True)
