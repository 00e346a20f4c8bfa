#!/bin/bash
# quick GPU round: parity tests + bench cfg2 (+ optional cfg3 with CFG3=1)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.err
cat gpurun_out/${TAG}_bench_cfg2.json
if [ "${CFG3:-0}" = "1" ]; then
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3.err
cat gpurun_out/${TAG}_bench_cfg3.json
fi
