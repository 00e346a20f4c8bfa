#!/bin/bash
# GPU parity tests + the two configs that do not fit one GPU's HBM as a table (table-free slab mode), one GPU.
TAG=${1:-big}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; tail -c 600 gpurun_out/${TAG}_bench_cfg4.err
cat gpurun_out/${TAG}_bench_cfg4.json
timeout 1200 python bench.py --workload cfg5 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; tail -c 600 gpurun_out/${TAG}_bench_cfg5.err
cat gpurun_out/${TAG}_bench_cfg5.json
