#!/bin/bash
python tools/sweep_chunks.py --chunks 0,3,4,5,6,7,8,9,10,12,14 --ring 32x8 > gpurun_out/r02_g_sweep_cfg2.txt 2>&1
python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0,2,3,4,5,6 --ring 32x8 > gpurun_out/r02_g_sweep_n200.txt 2>&1
python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0,1,2,3 --ring 32x8 > gpurun_out/r02_g_sweep_n500B.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x -k "counts" > gpurun_out/r02_g_tests.log 2>&1
cat gpurun_out/r02_g_sweep_cfg2.txt gpurun_out/r02_g_sweep_n200.txt gpurun_out/r02_g_sweep_n500B.txt; tail -3 gpurun_out/r02_g_tests.log
