#!/bin/bash
# round 2 experiment e: scan v5 timing (profile_count prints score_ms) + counting-kernel variants (G-form prefetch, 2-tree unroll)
python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 > gpurun_out/r02_e_scan_n500.log 2>&1
python tools/profile_count.py --n 100 --m 10000 --seed 2000 --reps 3 > gpurun_out/r02_e_scan_n100.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden_scores or scores_vs_oracle or table_free or shard or scan" > gpurun_out/r02_e_tests.log 2>&1
for lib in quartetscores_b200/libqscuda.so tools/variants/*.so; do
  echo "== $lib"
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --chunks 0 --ring 32x8 2>&1 | tail -1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0 --ring 32x8 2>&1 | tail -1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0 --ring 32x8 2>&1 | tail -1
done > gpurun_out/r02_e_count_variants.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qs_scan_kernel -s 1 -c 1 -o gpurun_out/r02_e_scan_n500 -f python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 2 > gpurun_out/r02_e_ncu.log 2>&1
cat gpurun_out/r02_e_scan_n500.log gpurun_out/r02_e_count_variants.txt; tail -3 gpurun_out/r02_e_tests.log
