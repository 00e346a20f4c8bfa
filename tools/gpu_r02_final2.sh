#!/bin/bash
# end-of-round validation after role Z / distance passes / AUTO-mode fix: what the driver runs + shard balance on one GPU + ncu evidence
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg2_1gpu.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 400 gpurun_out/${TAG}_bench_cfg2.err; cut -c1-300 gpurun_out/${TAG}_bench_cfg2_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg3_1gpu.json 2> gpurun_out/${TAG}_bench_cfg3.err; cut -c1-300 gpurun_out/${TAG}_bench_cfg3_1gpu.json
O=gpurun_out/${TAG}_shards.txt; : > $O
timeout 300 python tools/shard_costs.py --n 500 --m 1000 --G 8 >> $O 2>&1
timeout 300 python tools/shard_costs.py --n 1000 --m 500 --seed 4000 --p-missing 0 --p-contract 0 --G 8 >> $O 2>&1
cat $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qs_count_rows_kernel -s 1 -c 1 -o gpurun_out/${TAG}_count_rows_cfg2 -f python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 2 > gpurun_out/${TAG}_ncu_count.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference_cfg2.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference_cfg2.json
