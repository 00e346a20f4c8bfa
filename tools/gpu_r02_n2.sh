#!/bin/bash
# 2 GPUs: NCCL parity tests + the drop-in CLI with shards, cfg3 under the driver's torchrun line; then (GPU 0 only) ncu of the scan and distance kernels at cfg2
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_multi_gpu_nccl.py tests/test_cli_dropin.py -m gpu -q -x > gpurun_out/${TAG}_tests_n2.log 2>&1; tail -3 gpurun_out/${TAG}_tests_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg3_n2.json 2> gpurun_out/${TAG}_bench_cfg3_n2.err; cut -c1-330 gpurun_out/${TAG}_bench_cfg3_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:qs_scan_kernel -s 1 -c 1 -o gpurun_out/${TAG}_scan_cfg2 -f python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 2 > gpurun_out/${TAG}_ncu_scan.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:qs_dist_warp_kernel -s 1 -c 1 -o gpurun_out/${TAG}_dist_cfg2 -f python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 2 > gpurun_out/${TAG}_ncu_dist.log 2>&1
ls -la gpurun_out/*.ncu-rep
