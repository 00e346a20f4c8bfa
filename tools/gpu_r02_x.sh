#!/bin/bash
# role Z everywhere?  one shard (N = 1) with role Y (default) and with role Z for every d; and the bench's step-time jitter
TAG=${1:-r02_x}
mkdir -p gpurun_out
for i in 1 2 3; do timeout 250 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_wall_ms'], d['step_kernel_ms'])"; done 2>&1 | tee gpurun_out/${TAG}_jitter.txt
nproc; uptime
O=gpurun_out/${TAG}_z_everywhere.txt; : > $O
for cfg in "--n 500 --m 1000" "--n 100 --m 10000 --seed 2000" "--n 200 --m 4000 --seed 7"; do
  echo "default: $cfg" >> $O; timeout 300 python tools/shard_costs.py $cfg --G 1 >> $O 2>&1
  echo "all Z: $cfg" >> $O; QS_Z_MIN_BLOCKS=100000 QS_Z_MAX_SPAN=100000 timeout 300 python tools/shard_costs.py $cfg --G 1 >> $O 2>&1
done
echo "8 shards, Z for every shard (QS_Z_MIN_BLOCKS=100000 QS_Z_MAX_SPAN=100000)" >> $O
QS_Z_MIN_BLOCKS=100000 QS_Z_MAX_SPAN=100000 timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 8 >> $O 2>&1
cat $O
