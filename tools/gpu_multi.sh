#!/bin/bash
# multi-GPU bench under torchrun (the driver's launch line).  usage: tools/gpu_multi.sh TAG N [cfg3=0|1]
TAG=${1:-multi}; N=${2:-2}; CFG3=${3:-1}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_n$N.err; tail -c 800 gpurun_out/${TAG}_bench_cfg2_n$N.err
cat gpurun_out/${TAG}_bench_cfg2_n$N.json
if [ "$CFG3" = "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3_n$N.json 2> gpurun_out/${TAG}_bench_cfg3_n$N.err; tail -c 800 gpurun_out/${TAG}_bench_cfg3_n$N.err
cat gpurun_out/${TAG}_bench_cfg3_n$N.json
fi
