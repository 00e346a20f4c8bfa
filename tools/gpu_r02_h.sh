#!/bin/bash
python tools/sweep_chunks.py --chunks 0,3,4,5,6,7,8,10 --ring 32x8 > gpurun_out/r02_h_sweep_cfg2.txt 2>&1
python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0,2,3,4 --ring 32x8 > gpurun_out/r02_h_sweep_n200.txt 2>&1
python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0,1,2 --ring 32x8 > gpurun_out/r02_h_sweep_n500B.txt 2>&1
python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 > gpurun_out/r02_h_scan_n500.log 2>&1
python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 3 > gpurun_out/r02_h_scan_n100.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_h_tests.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qs_scan_kernel -s 1 -c 1 -o gpurun_out/r02_h_scan_n500 -f python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 2 > gpurun_out/r02_h_ncu.log 2>&1
cat gpurun_out/r02_h_sweep_cfg2.txt gpurun_out/r02_h_sweep_n200.txt gpurun_out/r02_h_sweep_n500B.txt gpurun_out/r02_h_scan_n500.log gpurun_out/r02_h_scan_n100.log; tail -3 gpurun_out/r02_h_tests.log
