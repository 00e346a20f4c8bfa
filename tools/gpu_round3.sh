#!/bin/bash
# tests + cfg2 + cfg3 + cfg4 (1 GPU) bench lines
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.err
cat gpurun_out/${TAG}_bench_cfg2.json
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3.err
cat gpurun_out/${TAG}_bench_cfg3.json
timeout 900 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; tail -c 600 gpurun_out/${TAG}_bench_cfg4.err
cat gpurun_out/${TAG}_bench_cfg4.json
