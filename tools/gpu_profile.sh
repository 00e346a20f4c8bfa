#!/bin/bash
# ncu captures of the counting kernels (one GPU).  Outputs under gpurun_out/<tag>_*.
TAG=${1:-prof}
mkdir -p gpurun_out
M="gpu__time_duration.sum"
# launch list of one bench step sequence (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# full capture: items kernel at cfg2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qs_count_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_rows100 python tools/profile_count.py --n 100 --m 10000 > gpurun_out/${TAG}_rows100.log 2>&1
# full capture: tiled kernel, class-B trees (missing taxa + polytomies), 300 taxa x 600 trees
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qs_count_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_rows300 python tools/profile_count.py --n 300 --m 600 --seed 3000 --p-missing 0.1 --p-contract 0.05 > gpurun_out/${TAG}_rows300.log 2>&1
ls -la gpurun_out/ | tail -8
