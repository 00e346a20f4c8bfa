#!/usr/bin/env python3
"""Host wall clock of every call of one step (cfg2 shape by default): where does a step's time go outside the kernels?"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100); ap.add_argument("--m", type=int, default=10000); ap.add_argument("--seed", type=int, default=2000)
ap.add_argument("--steps", type=int, default=6)
args = ap.parse_args()
from quartetscores_b200 import Context, QS_MODE_AUTO
from quartetscores_b200.computer import cint_bytes_for
from quartetscores_b200.newick import flatten_reference, parse_newick
from quartetscores_b200.synth import SyntheticInput
s = SyntheticInput(args.n, args.m, args.seed, k_max=20, want_newick=False)
ref = flatten_reference(parse_newick(s.ref_newick))
sync = torch.cuda.synchronize
with Context(args.n, cint_bytes_for(args.m), mode=QS_MODE_AUTO) as ctx:
    ctx.set_reference(ref)
    for step in range(args.steps):
        t = [time.perf_counter()]
        ctx.clear_trees(); ctx.add_trees(s.flat); t.append(time.perf_counter())
        sync(); t.append(time.perf_counter())
        ctx.count(); t.append(time.perf_counter())
        sync(); t.append(time.perf_counter())
        ctx.score(1); t.append(time.perf_counter())
        sync(); t.append(time.perf_counter())
        lt = ctx.last_timing(); t.append(time.perf_counter())
        names = ["add_trees", "sync", "count(call)", "sync", "score(call)", "sync", "last_timing"]
        print(f"step {step}: " + "  ".join(f"{nm}={1e3 * (b - a):.2f}" for nm, a, b in zip(names, t, t[1:])) + f"  kernels={lt}", flush=True)
    for step in range(3):
        t0 = time.perf_counter(); ctx.count(); ctx.score(1); sync(); t1 = time.perf_counter()
        print(f"resident step {step}: {1e3 * (t1 - t0):.2f} ms  kernels={ctx.last_timing()}", flush=True)
