#!/bin/bash
# role Z with tasks spanning consecutive d, Z for every range with < 8 whole d-blocks: parity, shard costs, bench
TAG=${1:-r02_y}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
O=gpurun_out/${TAG}_shards.txt; : > $O
timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 8 >> $O 2>&1
timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 1 >> $O 2>&1
timeout 300 python tools/shard_costs.py --n 100 --m 10000 --seed 2000 --G 8 >> $O 2>&1
echo "n=100 G=8 with role Y only where possible (QS_Z_MIN_BLOCKS=0)" >> $O
QS_Z_MIN_BLOCKS=0 timeout 300 python tools/shard_costs.py --n 100 --m 10000 --seed 2000 --G 8 >> $O 2>&1
timeout 300 python tools/shard_costs.py --n 100 --m 10000 --seed 2000 --G 1 >> $O 2>&1
timeout 300 python tools/shard_costs.py --n 1000 --m 500 --seed 4000 --p-missing 0 --p-contract 0 --G 8 >> $O 2>&1
cat $O
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_1gpu_$i.json 2> gpurun_out/${TAG}_bench_cfg2.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_cfg2_1gpu_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['step_wall_ms'], d['step_kernel_ms'], d['clocks'])"; done
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_1gpu.json 2> gpurun_out/${TAG}_bench_cfg3.err; cut -c1-300 gpurun_out/${TAG}_bench_cfg3_1gpu.json
