"""The multi-process NCCL path on real GPUs (needs >= 2 devices: skipped on the driver's single-GPU test box, run by
tools/gpu_r02_multi.sh under `gpurun --gpus N`): bench.py under the driver's torchrun line with N ranks must print the same
score checksum as the single-process run and match the reference digest of BASELINE configs[0]."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(n_ranks, workload):
    args = ["bench.py", "--gpus", str(n_ranks), "--workload", workload, "--steps", "2", "--warmup", "3", "--no-cpu-baseline"]
    if n_ranks > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}", "--master-addr", "127.0.0.1", "--master-port", str(29600 + n_ranks)] + args
    else:
        cmd = [sys.executable] + args
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("workload", ["cfg1", "cfg2"])
def test_n_rank_nccl_run_reproduces_the_single_gpu_scores(workload):
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    one = _bench(1, workload)
    assert one["golden"]["scores_within_1e-9"] is True
    for n in sorted({2, min(n_dev, 8)}):
        many = _bench(n, workload)
        assert many["n_gpus"] == n
        assert many["scores_sha256"] == one["scores_sha256"], f"N = {n} scores differ from N = 1"
        assert many["golden"]["scores_within_1e-9"] is True
