#!/bin/bash
python tools/sweep_chunks.py --chunks 0 --ring 32x2,16x4,10x3,8x4,8x8,5x4,4x8,3x8 > gpurun_out/r02_u_ring_cfg2.txt 2>&1
python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0 --ring 32x2,16x4,8x4,8x8,4x8 > gpurun_out/r02_u_ring_n200.txt 2>&1
python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0 --ring 32x2,16x4,8x8,4x8 > gpurun_out/r02_u_ring_n500B.txt 2>&1
cat gpurun_out/r02_u_ring_cfg2.txt gpurun_out/r02_u_ring_n200.txt gpurun_out/r02_u_ring_n500B.txt
