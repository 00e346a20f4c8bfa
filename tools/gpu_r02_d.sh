#!/bin/bash
# distance kernel with the matrix in shared memory: parity, then kernel times with and without it; ncu of the scan and distance kernels at cfg2
TAG=${1:-r02_d}
mkdir -p gpurun_out
QS_DIST_SMEM_MATRIX=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "dist or golden or malformed or class or unary or deep or cfg1 or cfg2" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
echo "QS_DIST_SMEM_MATRIX=1"; QS_DIST_SMEM_MATRIX=1 timeout 120 python tools/step_timing.py --steps 4 2>&1 | grep "resident step" | tail -2
echo "QS_DIST_SMEM_MATRIX=0"; QS_DIST_SMEM_MATRIX=0 timeout 120 python tools/step_timing.py --steps 4 2>&1 | grep "resident step" | tail -2
timeout 250 ncu --set full --clock-control none --import-source on -k regex:qs_scan_kernel -s 1 -c 1 -o gpurun_out/${TAG}_scan_cfg2 -f python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 2 > gpurun_out/${TAG}_ncu_scan.log 2>&1
QS_DIST_SMEM_MATRIX=1 timeout 250 ncu --set full --clock-control none --import-source on -k regex:qs_dist_warp_kernel -s 1 -c 1 -o gpurun_out/${TAG}_dist_cfg2 -f python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 2 > gpurun_out/${TAG}_ncu_dist.log 2>&1
ls -la gpurun_out/${TAG}*.ncu-rep
