#!/bin/bash
# 8-GPU runs of the two scaling configs (BASELINE configs[2], configs[3]) under the driver's torchrun line.
TAG=${1:-m8}; N=${2:-8}
mkdir -p gpurun_out
for W in cfg3 cfg4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${W}_n$N.json 2> gpurun_out/${TAG}_bench_${W}_n$N.err; tail -c 500 gpurun_out/${TAG}_bench_${W}_n$N.err
cat gpurun_out/${TAG}_bench_${W}_n$N.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_n$N.err
cat gpurun_out/${TAG}_bench_cfg2_n$N.json
