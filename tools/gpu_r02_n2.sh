#!/bin/bash
# 2 GPUs: NCCL parity tests + the drop-in CLI with shards, cfg3 under the driver's torchrun line
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_multi_gpu_nccl.py tests/test_cli_dropin.py -m gpu -q -x > gpurun_out/${TAG}_tests_n2.log 2>&1; tail -3 gpurun_out/${TAG}_tests_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg3_n2.json 2> gpurun_out/${TAG}_bench_cfg3_n2.err; cut -c1-330 gpurun_out/${TAG}_bench_cfg3_n2.json
