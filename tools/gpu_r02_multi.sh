#!/bin/bash
# round 2 multi-GPU run: usage tools/gpu_r02_multi.sh TAG N   (gpurun --gpus N)
TAG=${1:-r02_multi}; N=${2:-2}; TESTS=${3:-1}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
if [ "$TESTS" = "1" ]; then timeout 900 python -m pytest tests/test_multi_gpu_nccl.py tests/test_cli_dropin.py -m gpu -q -x > gpurun_out/${TAG}_tests_n$N.log 2>&1; tail -3 gpurun_out/${TAG}_tests_n$N.log; fi
# the driver's launch line: default workload under torchrun = cfg3
for n in $N; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg3_n$n.json 2> gpurun_out/${TAG}_bench_cfg3_n$n.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3_n$n.err; cat gpurun_out/${TAG}_bench_cfg3_n$n.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2_n$N.json 2> gpurun_out/${TAG}_bench_cfg2_n$N.err; cat gpurun_out/${TAG}_bench_cfg2_n$N.json
if [ "$N" -ge 4 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4_n$N.json 2> gpurun_out/${TAG}_bench_cfg4_n$N.err; tail -c 600 gpurun_out/${TAG}_bench_cfg4_n$N.err; cat gpurun_out/${TAG}_bench_cfg4_n$N.json
fi
