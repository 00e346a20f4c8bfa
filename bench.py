#!/usr/bin/env python3
"""bench.py — headline benchmark of the quartet counting/scoring hot path (BASELINE.json metric:
quartet x tree evaluations per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg1|cfg2|cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole synthetic workload: build all gene-tree
distance matrices, count every quartet x tree, scan the table into LQ-IC/QP-IC/EQP-IC.
  value : step with the flattened gene trees already resident in HBM (qs_count + scoring), CUDA events
  e2e   : the same through the C ABI from pinned HOST buffers: qs_clear_trees + qs_add_trees (H2D) +
          qs_count + qs_score (D2H of the scores), every step
N > 1: one process per GPU, each owns a balanced range of the quartet rank space (outer index d); the
only communication is one all-reduce(min) of per-edge LQ-IC partials and one all-reduce(sum) of the
per-node-pair topology sums (NCCL).  The job is fixed as N grows: "scaling": "strong".

--impl reference times the UNMODIFIED reference binary (oracle/_ref/QuartetScores, built from
/root/reference by oracle/Makefile) with -t <all host cores> on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import gc
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from math import comb

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0..2]; seeds follow SURVEY.md §8d (1000*config + replicate)
    "cfg1": dict(n_taxa=50, n_trees=1000, seed=1000, k_max=10, label="50 taxa x 1,000 NNI/SPR-perturbed gene trees"),
    "cfg2": dict(n_taxa=100, n_trees=10000, seed=2000, k_max=20, label="100 taxa x 10,000 gene trees, full uint16 lookup table (3.9M quartets) on 1 B200"),
    "cfg3": dict(n_taxa=500, n_trees=5000, seed=3000, k_max=20, p_missing=0.1, p_contract=0.05,
                 label="500 taxa x 5,000 gene trees with missing taxa and multifurcations, uint16 table (15.4 GB)"),
    # configs[3..4]: the 248.5 GB / 3.99 TB tables do not fit one GPU -> table-free (slab-streamed) mode unless the shard's table fits
    "cfg4": dict(n_taxa=1000, n_trees=2000, seed=4000, k_max=20,
                 label="1,000 taxa x 2,000 gene trees, 248.5 GB uint16 table: sharded in HBM when a shard fits, else table-free slabs"),
    "cfg5": dict(n_taxa=2000, n_trees=1000, seed=5000, k_max=20,
                 label="-s savemem mode, 2,000 taxa x 1,000 gene trees, table-free: counts evaluated on the fly, never stored"),
}
TABLE_BYTES_LIMIT = 120e9      # a shard's uint16 table above this is not kept resident (180 GB HBM3e minus matrices and slack)
# one packed fp16x2 compare (HSET2 lane-op) decides one topology slot of TWO quartets.  A gene tree that
# resolves every quartet (class A) needs 2 slots per quartet x tree = 1.0 HSET2 lane-op per evaluation, any other
# tree (class B) needs 3 = 1.5.
HSET2_LANEOPS_PER_EVAL = {"A": 1.0, "B": 1.5}
INT32_LANEOPS_PER_EVAL = 9.0       # 3 adds + 3 compares + 3 predicated increments
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the counting kernel from the committed `ncu --set full`
# capture of the same workload on one GPU (profiles/r02_final_count_rows_cfg2_ncu_full.txt: 244.6 MB read + 36.2 MB written; the algorithmic
# bytes are 208 MB of matrices + 23.5 MB of table); it is NOT measured by the run that prints it (the entry says so); null where no
# capture of that exact workload exists
NCU_DRAM_TRAFFIC = {("cfg2", 1): {"bytes": 244608000 + 36208128, "source": "profiles/r02_final_count_rows_cfg2_ncu_full.txt", "not_this_run": True}}


def hbm_peak_gbs():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0      # B200_PROFILING.md fallback ("of fallback")


def make_input(w, want_newick=False):
    from quartetscores_b200.synth import SyntheticInput

    kw = {k: v for k, v in w.items() if k not in ("label",)}
    return SyntheticInput(want_newick=want_newick, **kw)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int, interval_ms: int = 20):
        # the poll itself costs the sampled GPU a few per cent while it runs (rank 0's counting kernel was 4-7 % slower than the other
        # ranks' with 50 polls per second): long timed regions are polled less often
        self.idx, self.rows, self.proc, self.interval_ms = gpu_index, [], None, int(max(20, min(500, interval_ms)))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.interval_ms), "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.05 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# -------------------------------------------------------------------------------------------------
# reference arm: the unmodified CPU implementation
# -------------------------------------------------------------------------------------------------

def run_reference_binary(ref_nwk, eval_lines, threads, tmp, tag):
    """Run oracle/_ref/QuartetScores once; returns (elapsed_s, counting_s, scoring_s, table_kind)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "QuartetScores")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    rp, ep, op = (os.path.join(tmp, f"{tag}.{x}") for x in ("ref.nwk", "eval.nwk", "out.nwk"))
    with open(rp, "w") as f:
        f.write(ref_nwk + "\n")
    with open(ep, "w") as f:
        f.write("\n".join(eval_lines) + "\n")
    if os.path.exists(op):
        os.remove(op)   # the reference refuses to overwrite (src/QuartetScores.cpp:81-85)
    out = subprocess.run([exe, "-r", rp, "-e", ep, "-o", op, "-t", str(threads)], capture_output=True, text=True, check=True).stdout
    took = [int(x) for x in re.findall(r"It took: (\d+) microseconds", out)]
    elapsed = int(re.search(r"Elapsed time: (\d+) microseconds", out).group(1))
    kind = "fast" if "Using runtime-efficient" in out else "compact"
    return elapsed * 1e-6, took[0] * 1e-6, took[1] * 1e-6, kind


def reference_full_estimate(nq, m, sample, elapsed, count_s, score_s):
    """Evaluations/s of the reference on the WHOLE workload, from a run on the first `sample` trees: counting and Newick
    parsing grow linearly with the tree count, the scoring pass over the C(n,4) quartets does not (it is the same for any
    m), so quoting the sample's own throughput would understate the reference."""
    r = m / float(sample)
    other = max(0.0, elapsed - count_s - score_s)
    full_s = count_s * r + score_s + other * r
    return nq * m / full_s, full_s


def reference_sample_size(w, target_evals):
    nq = comb(w["n_taxa"], 4)
    s = int(target_evals / nq)
    return max(256, min(w["n_trees"], s))    # >= 256 trees keeps CINT = uint16 (src/QuartetScores.cpp:115-123)


REFERENCE_MAX_QUARTETS = 5e7      # the reference's scoring pass visits every quartet under a critical section (5.8 s for 3.9e6 at cfg2):
                                   # at 500 taxa (2.6e9 quartets) one run takes hours whatever the tree sample, so it is not attempted


def reference_infeasible(w):
    nq = comb(w["n_taxa"], 4)
    if nq <= REFERENCE_MAX_QUARTETS:
        return None
    return (f"the reference binary is not run on this workload: its scoring pass is serialised over all {nq:.3g} quartets "
            f"(about {5.8 * nq / 3.9e6 / 3600:.1f} h at the cfg2 rate) and no tree sample shortens it; cfg2 is the workload both arms can run")


def reference_arm(args, w, wname):
    """The UNMODIFIED reference binary on the WHOLE workload (no tree sample, no extrapolation): one warm-up run and
    min(steps, 2) timed runs -- a cfg2 run takes about a minute on 16 cores, so --steps/--warmup beyond that are ignored."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    why = reference_infeasible(w)
    if why:
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)
        return 0
    cores = os.cpu_count() or 1
    m = w["n_trees"]
    inp = make_input(w, want_newick=True)
    nq = comb(w["n_taxa"], 4)
    n_warm, n_timed = min(args.warmup, 1), max(1, min(args.steps, 2))
    times, parts = [], []
    kind = "?"
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(n_warm + n_timed):
            el, cnt, sc, kind = run_reference_binary(inp.ref_newick, inp.eval_newick, cores, tmp, "r")
            if i >= n_warm:
                times.append(el)
                parts.append((cnt, sc))
    mean = lambda xs: sum(xs) / len(xs)
    value = nq * m / mean(times)
    line = {
        "impl": "reference", "metric": "quartet_tree_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": n_timed,
        "warmup": n_warm, "ms_per_step": 1e3 * mean(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": {"workload": f"{wname}: {w['label']}", "seed": w["seed"], "quartets": nq, "trees": m,
                   "note": f"whole workload every step; {n_warm} warm-up + {n_timed} timed runs (a run takes about a minute: --steps/--warmup beyond that are ignored)"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "reference",
                         "sample": f"oracle/_ref/QuartetScores -t {cores}, {kind} table, ALL {m} gene trees (no sample, no extrapolation): {mean(times):.2f} s end-to-end "
                                   f"({mean([p[0] for p in parts]):.2f} s counting, {mean([p[1] for p in parts]):.2f} s scoring), Newick files in, annotated Newick out"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_cli": {"seconds": mean(times), "what": f"oracle/_ref/QuartetScores -r ref.nwk -e eval.nwk -o out.nwk -t {cores}: the binary's own 'Elapsed time'"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------

GOLDEN_BIG = {"cfg1": "cfg1_50x1000", "cfg2": "cfg2_100x10000"}


def _load_golden(wname, w):
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "big", GOLDEN_BIG.get(wname, "-") + ".npz")
    if not os.path.exists(path):
        return None, None
    z = np.load(path)
    spec = json.loads(z["spec"].item())
    if any(spec.get(k) != w.get(k) for k in ("n_taxa", "n_trees", "seed", "k_max")):
        return None, None
    return z, path


def golden_check(wname, w, scores):
    """Compare this run's scores with the unmodified reference's on the same input (committed digest), where one exists."""
    import numpy as np

    z, path = _load_golden(wname, w)
    if z is None:
        return None
    worst, same_inf = 0.0, True
    for got, key in zip(scores, ("lqic", "qpic", "eqpic")):
        want = z[key]
        same_inf = same_inf and bool(np.array_equal(np.isinf(got), np.isinf(want)))
        fin = np.isfinite(want) & np.isfinite(got)
        if fin.any():
            worst = max(worst, float(np.abs(got[fin] - want[fin]).max()))
    return {"file": os.path.relpath(path, ROOT), "reference": "oracle/_ref/qs_ref_dump (unmodified reference headers), fast table",
            "scores_within_1e-9": bool(same_inf and worst <= 1e-9), "max_abs_diff": worst}


def cli_end_to_end(w, wname):
    """End-to-end seconds of the drop-in CLI: the reference's own main linked against libqscuda.so
    (oracle/_ref/QuartetScoresB200, integration/main_b200.cpp): Newick files in, annotated Newick out, process start and
    CUDA context creation included.  The reference binary's number for the same files is the reference arm's e2e_cli."""
    exe = os.path.join(ROOT, "oracle", "_ref", "QuartetScoresB200")
    if not os.path.exists(exe):
        return {"seconds": None, "what": "oracle/_ref/QuartetScoresB200 not built (needs /root/reference at build time)"}
    try:
        inp = make_input(w, want_newick=True)
        z, _ = _load_golden(wname, w)
        runs = []
        with tempfile.TemporaryDirectory() as tmp:
            rp, ep = os.path.join(tmp, "ref.nwk"), os.path.join(tmp, "eval.nwk")
            inp.write(rp, ep)
            same = None
            for i in range(3):
                op = os.path.join(tmp, f"out{i}.nwk")
                t0 = time.time()
                out = subprocess.run([exe, "-r", rp, "-e", ep, "-o", op, "-t", str(os.cpu_count() or 1)], capture_output=True, text=True, check=True).stdout
                wall = time.time() - t0
                el = re.search(r"Elapsed time: (\d+) microseconds", out)
                runs.append({"wall_s": round(wall, 4), "elapsed_s": int(el.group(1)) * 1e-6 if el else None})
                if z is not None:
                    same = open(op).read() == z["out_newick"].item()
        return {"seconds": min(r["wall_s"] for r in runs), "runs": runs, "output_identical_to_reference": same,
                "what": "wall clock of `oracle/_ref/QuartetScoresB200 -r ref.nwk -e eval.nwk -o out.nwk` (fork to exit: process start, CUDA context, "
                        "Newick ingest, counting, scoring, writing); best of 3, every run listed; elapsed_s is the binary's own 'Elapsed time' line"}
    except Exception as e:
        return {"seconds": None, "what": f"failed: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (long workloads: cfg4/cfg5 exploration runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed region (the line then has no clocks evidence)")
    args = ap.parse_args()
    if args.impl == "ours" and not args.no_e2e:
        args.warmup = max(args.warmup, 3)       # timing rule: W >= 3 (exploration runs with --no-e2e may use fewer)

    # Default workload: the plain single-process run (the driver's BENCH, N = 1) times BASELINE configs[1] (cfg2), the one
    # configuration the reference binary also runs in minutes, so both arms share it.  A run launched through
    # torch.distributed.run (the driver's 1/2/4/8 scaling runs, N = 1 included) times configs[2] (cfg3), the configuration
    # BASELINE.json lists "at 1/2/4/8 B200": a 5 ms cfg2 step cannot show how rank-space shards scale.
    under_torchrun = "TORCHELASTIC_RUN_ID" in os.environ or "WORLD_SIZE" in os.environ
    wname = args.workload or ("cfg3" if under_torchrun else "cfg2")
    w = WORKLOADS[wname]
    if args.impl == "reference":
        try:
            return reference_arm(args, w, wname)
        except FileNotFoundError as e:
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": f"reference binary not built ({e}); run `make -C oracle ref` where /root/reference exists"}))
            return 0

    import numpy as np
    import torch

    from quartetscores_b200 import QS_MODE_AUTO, QS_MODE_TABLE_FREE, Context
    from quartetscores_b200.computer import cint_bytes_for
    from quartetscores_b200.multi import shard_bounds
    from quartetscores_b200.multi import score_distributed
    from quartetscores_b200.newick import flatten_reference, parse_newick

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quartetscores_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

        # NCCL prints its version banner on stdout when the communicator is created; rank 0's stdout carries ONE JSON line, so
        # stdout points at stderr until the first collective is through
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    n, m = w["n_taxa"], w["n_trees"]
    nq = comb(n, 4)
    inp = make_input(w)
    ref = flatten_reference(parse_newick(inp.ref_newick))
    flat = inp.flat
    # pinned host buffers: the e2e leg copies from these every step
    h_off = torch.from_numpy(np.ascontiguousarray(flat.node_offsets)).pin_memory()
    h_par = torch.from_numpy(np.ascontiguousarray(flat.parent)).pin_memory()
    h_leaf = torch.from_numpy(np.ascontiguousarray(flat.leaf_lookup_id)).pin_memory()

    # cfg5 is the reference's -s run: no resident table by request; everything else keeps its shard's table in HBM when it fits
    # beside the distance matrices and falls back to table-free slabs when it does not (QS_MODE_AUTO: cfg4 on 1-2 GPUs)
    ctx = Context(n, cint_bytes_for(m), mode=QS_MODE_TABLE_FREE if wname == "cfg5" else QS_MODE_AUTO, device=local_rank, shard_index=rank, shard_count=world)
    if wname == "cfg5":
        ctx.set_count_scale(2)          # the reference's -s table semantics (doubled, CINT-wrapped counts)
    score_scale = 2 if wname == "cfg5" else 1
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_reference(ref)
    E = ref.edge_count

    def add_trees():
        ctx.clear_trees()
        ctx.add_trees_ptr(flat.n_trees, h_off.data_ptr(), h_par.data_ptr(), h_leaf.data_ptr())

    def score():
        return score_distributed(ctx, score_scale)     # 1 GPU: qs_score; N GPUs: partial scan + all-reduce(min) + all-reduce(sum) + finalise

    def step_resident():
        ctx.count()
        return score()

    def step_e2e():
        add_trees()
        ctx.count()
        return score()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def host_state():
        # who else used the host during the timed region: context switches of this process and the box's CPU time (jiffies) by kind
        st = {}
        try:
            for line in open("/proc/self/status"):
                if line.startswith(("voluntary_ctxt_switches", "nonvoluntary_ctxt_switches")):
                    st[line.split(":")[0]] = int(line.split()[1])
            f = open("/proc/stat").readline().split()
            st.update(dict(zip(("user", "nice", "system", "idle", "iowait", "irq", "softirq", "steal"), (int(x) for x in f[1:9]))))
        except Exception:
            pass
        return st

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()
        hs0 = host_state()
        t0 = time.time()
        e0.record(stream)
        kernel_ms = []
        for _ in range(k):
            ts = time.perf_counter()
            out = fn()
            kernel_ms.append(ctx.last_timing())
            kernel_ms[-1]["host_wall_ms"] = 1e3 * (time.perf_counter() - ts)      # diagnostic: this step on the host clock (last_timing waits for the step's events)
        e1.record(stream)
        barrier()
        t1 = time.time()
        hs1 = host_state()
        gc.enable()
        host_delta.clear()
        host_delta.update({k_: hs1[k_] - hs0[k_] for k_ in hs0 if k_ in hs1})
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, kernel_ms, (t0, t1)

    host_delta = {}
    add_trees()
    ctx.rebalance_shards()       # shard ranges for the class mix of these trees (same trees on every rank -> same ranges, no exchange)
    t_w = time.time()
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    step_s = (time.time() - t_w) / max(1, args.warmup)
    table_free = not ctx.table_resident()
    sampler = ClockSampler(local_rank, interval_ms=int(1e3 * step_s * args.steps / 25))      # ~25 samples over the timed region, at least every 500 ms
    if rank == 0 and not args.no_clocks:
        # nvidia-smi takes a few hundred ms to come up and holds driver locks while it does: the timed region starts only after its
        # first sample has arrived; untimed steps keep the GPU busy meanwhile
        sampler.start()
        t_s = time.time()
        while not sampler.rows and time.time() - t_s < 5.0:
            if dist is None:
                step_resident()
            else:
                time.sleep(0.01)           # (steps hold collectives: the other ranks wait at the timed region's first barrier instead)
        torch.cuda.synchronize()
    l0 = ctx.launch_count()
    ms_res, scores, kt, span = timed(step_resident, args.steps)
    host_res = dict(host_delta)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop(*span) if rank == 0 else None
    if args.no_e2e:
        ms_e2e = None
    else:
        for _ in range(2):
            step_e2e()
        ms_e2e, scores_e2e, _, _ = timed(step_e2e, args.steps)
        assert all(np.array_equal(a, b) for a, b in zip(scores, scores_e2e)), "resident and e2e legs disagree"

    # roofline of the dominant kernel (counting): algorithmic HSET2 lane-ops / measured kernel time vs the live-measured
    # HSET2 issue rate of this GPU (the pipe the kernel is bound by)
    hset2_peak, int32_peak = ctx.measure_alu_peak()
    count_ms = statistics.mean(t["count_ms"] for t in kt)
    r0, r1 = ctx.shard_range()
    nA, nB = ctx.tree_classes()
    my_quartets = r1 - r0
    my_evals = my_quartets * m
    algo_laneops = my_quartets * (nA * HSET2_LANEOPS_PER_EVAL["A"] + nB * HSET2_LANEOPS_PER_EVAL["B"])
    per_rank_ms = None
    if dist is not None:      # the counting kernel's time on every rank: shows how even the rank-space shards are
        g = [torch.zeros(3, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(g, torch.tensor([count_ms, statistics.mean(t["dist_ms"] for t in kt), statistics.mean(t["score_ms"] for t in kt)], device="cuda", dtype=torch.float64))
        per_rank_ms = [[round(float(v), 3) for v in x.tolist()] for x in g]
    achieved = algo_laneops / (count_ms * 1e-3)
    roofline = {
        "bound": "alu_issue", "achieved": achieved / 1e12, "peak": hset2_peak / 1e12, "unit": "Tlaneop/s", "frac": achieved / hset2_peak,
        "traffic": NCU_DRAM_TRAFFIC.get((wname, world)),      # ncu DRAM bytes of one launch from a committed capture (never measured by this run), else null
        "kernel": "qs_count_rows_kernel", "kernel_ms": count_ms,
        "dist_kernel_ms": statistics.mean(t["dist_ms"] for t in kt), "score_kernel_ms": statistics.mean(t["score_ms"] for t in kt),
        "peak_source": "measured live on this GPU: HSET2 (fp16x2 compare -> mask) lane-op rate in the kernel's own 2xHSET2+IADD3 mix (qs_measure_alu_peak)",
        "enumeration_efficiency": None,
        "mix_ceiling": {"frac_of_peak": 0.855, "frac": achieved / (0.855 * hset2_peak),
                        "note": "the kernel pairs every HSET2 (ALU pipe) with an IMAD.IADD (FMA pipe); that pair issues at 3.42 of 4 warp-instr/clk/SM in isolation "
                                "(tools/ubench_mix2.cu, profiles/r01_w_ubench_mix2.txt), i.e. 0.855 of the HSET2 peak is the most this instruction mix can reach"},
        "algorithmic_laneops_per_eval": algo_laneops / my_evals,
        "tree_classes": {"A_fully_resolved": nA, "B_general": nB},
        "per_rank_ms[count,dist,score]": per_rank_ms,
        "int32_equiv": {"achieved": INT32_LANEOPS_PER_EVAL * my_evals / (count_ms * 1e-3) / 1e12, "peak": int32_peak / 1e12, "unit": "Tlaneop/s",
                        "frac": INT32_LANEOPS_PER_EVAL * my_evals / (count_ms * 1e-3) / int32_peak,
                        "note": "SURVEY §8d accounting: 9 scalar int32 lane-ops per evaluation vs the measured INT32 (LOP3/IADD3) lane rate"},
        "scan": None if table_free else {
            "kernel": "qs_scan_kernel", "bound": "hbm", "kernel_ms": statistics.mean(t["score_ms"] for t in kt),
            "algorithmic_bytes": (r1 - r0) * 3 * cint_bytes_for(m), "achieved_gbs": (r1 - r0) * 3 * cint_bytes_for(m) / (statistics.mean(t["score_ms"] for t in kt) * 1e-3) / 1e9,
            "peak_gbs": hbm_peak_gbs(), "frac": (r1 - r0) * 3 * cint_bytes_for(m) / (statistics.mean(t["score_ms"] for t in kt) * 1e-3) / 1e9 / hbm_peak_gbs(),
            "note": "the table scan reads every entry of this rank's table once (3 x CINT bytes per quartet); score_kernel_ms also holds the clearing of the per-pair arrays"},
        "hbm": {"algorithmic_bytes": (r1 - r0) * 6 + 2 * n * n * m, "achieved_gbs": ((r1 - r0) * 6 + 2 * n * n * m) / (count_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak_gbs(),
                "note": "table written once + distance matrices read once; the path is ALU-bound by three orders of magnitude (traffic = ncu DRAM bytes of one launch)"},
    }

    # checksum of the three score vectors (bit patterns): must be the same at every N (integer sums and exact minima do not
    # depend on how the rank space is sharded), and for cfg1/cfg2 the scores are compared with the unmodified reference's
    # (tests/golden/big/*.npz, generated by tests/golden/make_golden_big.py from oracle/_ref)
    scores_sha256 = hashlib.sha256(b"".join(np.ascontiguousarray(x, np.float64).tobytes() for x in scores)).hexdigest()
    golden = golden_check(wname, w, scores)
    cli = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not reference_infeasible(w):
        cli = cli_end_to_end(w, wname)

    line = {
        "metric": "quartet_tree_evals_per_s", "value": nq * m * args.steps / (ms_res * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16x2 compares of exact small integers -> u16x2 integer counters (HSET2 mask + IMAD.IADD), u16 table, f64 scores", "data": "synthetic",
        "config": {"workload": f"{wname}: {w['label']}", "seed": w["seed"], "quartets": nq, "trees": m,
                   "l2": "inputs larger than L2: the distance matrices (%.0f MB) are rebuilt and re-streamed every step" % (2e-6 * n * ((n + 7) // 8 * 8) * m),
                   "parallelism": f"rank-space shards x{world}", "table": "table-free slabs (counted, scanned, discarded)" if table_free else "resident in HBM"},
        "e2e": None if args.no_e2e else {
            "value": nq * m * args.steps / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
            "h2d_bytes_per_step": int(h_off.numel() * 8 + h_par.numel() * 4 + h_leaf.numel() * 4), "d2h_bytes_per_step": int(3 * E * 8)},
        "gpu_launches": int(launches),
        "scores_sha256": scores_sha256,
        "golden": golden,
        "e2e_cli": cli,
        "clocks": clocks,
        "host": host_res,          # context switches of this process and CPU jiffies of the box over the timed (resident) region
        "step_wall_ms": [round(t["host_wall_ms"], 3) for t in kt],          # each timed step on the host clock and the sum of its three kernel groups:
        "step_kernel_ms": [round(t["dist_ms"] + t["count_ms"] + t["score_ms"], 3) for t in kt],      # a gap between the two is host time (launches, jitter), not GPU work
        "roofline": roofline,
    }

    if rank == 0 and not args.no_cpu_baseline and reference_infeasible(w):
        line["cpu_baseline"] = {"value": None, "unit": "evals/s", "cores": 0, "kind": "reference", "sample": reference_infeasible(w)}
    elif rank == 0 and not args.no_cpu_baseline:
        # reference binary on this box's host cores, bounded sample of the same workload
        try:
            cores = os.cpu_count() or 1
            sample = reference_sample_size(w, 6.0e9)
            sub = make_input(dict(w, n_trees=sample), want_newick=True)
            with tempfile.TemporaryDirectory() as tmp:
                el, cnt, sc, kind = run_reference_binary(sub.ref_newick, sub.eval_newick, cores, tmp, "cpu")
            full_value, full_s = reference_full_estimate(nq, m, sample, el, cnt, sc)
            line["cpu_baseline"] = {"value": full_value, "unit": "evals/s", "cores": cores, "kind": "reference", "sample_value": nq * sample / el,
                                    "sample": f"oracle/_ref/QuartetScores -t {cores}, {kind} table, first {sample} of {m} trees: {el:.2f} s end-to-end ({cnt:.2f} s counting, "
                                              f"{sc:.2f} s scoring); value = whole-workload estimate ({full_s:.1f} s: counting and parsing scaled by the tree count, scoring "
                                              f"constant); the sample's own throughput is sample_value"}
        except Exception as e:  # the oracle port is the documented fallback
            line["cpu_baseline"] = {"value": None, "unit": "evals/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
