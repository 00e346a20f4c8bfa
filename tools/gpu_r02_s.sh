#!/bin/bash
# scan kernel with per-warp candidate queues: parity, then kernel times at cfg2 and cfg3
TAG=${1:-r02_s2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 120 python tools/step_timing.py --steps 4 2>&1 | grep "resident step" | tail -2
timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 1 2>&1 | tail -1
timeout 300 python tools/shard_costs.py --n 200 --m 3000 --seed 7 --G 1 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_1gpu.json 2> gpurun_out/${TAG}_bench_cfg2.err; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_cfg2_1gpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['dist_kernel_ms'], d['roofline']['score_kernel_ms'], d['golden'])"
