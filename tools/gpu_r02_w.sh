#!/bin/bash
# role Z for the d of cut d-blocks + warp-parallel distance passes: parity first, then shard costs and the bench
TAG=${1:-r02_w}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for zb in 0 2 4 9; do
  echo "QS_Z_MIN_BLOCKS=$zb" >> gpurun_out/${TAG}_shards_n500B.txt
  QS_Z_MIN_BLOCKS=$zb timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 8 >> gpurun_out/${TAG}_shards_n500B.txt 2>&1
done
timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 1 >> gpurun_out/${TAG}_shards_n500B.txt 2>&1; cat gpurun_out/${TAG}_shards_n500B.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_1gpu.json 2> gpurun_out/${TAG}_bench_cfg2.err; cut -c1-300 gpurun_out/${TAG}_bench_cfg2_1gpu.json
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_1gpu.json 2> gpurun_out/${TAG}_bench_cfg3.err; cut -c1-300 gpurun_out/${TAG}_bench_cfg3_1gpu.json
