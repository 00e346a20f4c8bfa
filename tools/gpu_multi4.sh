#!/bin/bash
TAG=${1:-m4}; N=${2:-4}
mkdir -p gpurun_out
for W in cfg3 cfg4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${W}_n$N.json 2> gpurun_out/${TAG}_bench_${W}_n$N.err; tail -c 300 gpurun_out/${TAG}_bench_${W}_n$N.err
cat gpurun_out/${TAG}_bench_${W}_n$N.json
done
