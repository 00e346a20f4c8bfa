#!/usr/bin/env python3
"""Which call of a step puts the calling thread to sleep (voluntary context switches per call)?  A thread that sleeps waiting
for the GPU can be woken late on a busy virtualised host; a thread that spins cannot."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quartetscores_b200 import Context, QS_MODE_AUTO
from quartetscores_b200.computer import cint_bytes_for
from quartetscores_b200.newick import flatten_reference, parse_newick
from quartetscores_b200.synth import SyntheticInput

def vol():
    for line in open("/proc/thread-self/status"):
        if line.startswith("voluntary_ctxt_switches"):
            return int(line.split()[1])

s = SyntheticInput(100, 10000, 2000, k_max=20, want_newick=False)
ref = flatten_reference(parse_newick(s.ref_newick))
use_torch_stream = len(sys.argv) > 1 and sys.argv[1] == "torch"
with Context(100, cint_bytes_for(10000), mode=QS_MODE_AUTO) as ctx:
    if use_torch_stream:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_reference(ref); ctx.add_trees(s.flat)
    for step in range(6):
        row = []
        for name, fn in (("clear+add", lambda: (ctx.clear_trees(), ctx.add_trees(s.flat))), ("count", ctx.count), ("score", lambda: ctx.score(1)),
                         ("last_timing", ctx.last_timing), ("torch.sync", torch.cuda.synchronize)):
            v0, t0 = vol(), time.perf_counter(); fn(); row.append(f"{name}: {vol() - v0} sw {1e3 * (time.perf_counter() - t0):.2f} ms")
        print(f"step {step}  " + " | ".join(row), flush=True)
