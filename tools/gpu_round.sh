#!/bin/bash
# One gpurun call: GPU parity tests, microbenchmark, bench lines.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
if [ -x tools/ubench_pipes ]; then timeout 300 tools/ubench_pipes > gpurun_out/${TAG}_ubench.txt 2>&1; fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.err
cat gpurun_out/${TAG}_bench_cfg2.json
if [ "${CFG3:-1}" = "1" ]; then
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3.err
cat gpurun_out/${TAG}_bench_cfg3.json
fi
