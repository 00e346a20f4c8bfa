#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for n in 128 150 176; do for t in 512 256; do
  QS_CR_THREADS=$t python tools/sweep_chunks.py --n $n --m 3000 --seed 2100 --chunks 0 --ring 32x8 2>&1 | tail -1 | sed "s/^/threads=$t /"
done; done > gpurun_out/${TAG}_shape_threshold.txt 2>&1
cat gpurun_out/${TAG}_shape_threshold.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.err
cat gpurun_out/${TAG}_bench_cfg2.json
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3.err
cat gpurun_out/${TAG}_bench_cfg3.json
