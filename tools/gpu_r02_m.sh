#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x -k "score or table_free or shard or scan or golden" > gpurun_out/r02_o_tests.log 2>&1; tail -3 gpurun_out/r02_o_tests.log
for t in 352 512; do
QS_SCAN_THREADS=$t python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 2>&1 | tail -1
QS_SCAN_THREADS=$t python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 3 2>&1 | tail -1
done > gpurun_out/r02_o_scan.log 2>&1; cat gpurun_out/r02_o_scan.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qs_scan_kernel -s 1 -c 1 -o gpurun_out/r02_o_scan_n500 -f python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 2 > gpurun_out/r02_o_ncu.log 2>&1
