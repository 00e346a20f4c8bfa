#!/bin/bash
for lib in quartetscores_b200/libqscuda.so tools/variants/libqs_noxr.so tools/variants/libqs_runf.so; do
  echo "== $lib"
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --chunks 0,3,5,7,9,12 --ring 32x8 2>&1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0,1,2,3 --ring 32x8 2>&1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0 --ring 32x8 2>&1 | head -1
done > gpurun_out/r02_j_variants.txt 2>&1
python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 > gpurun_out/r02_j_scan_n500.log 2>&1
cat gpurun_out/r02_j_variants.txt gpurun_out/r02_j_scan_n500.log
