#!/bin/bash
# One gpurun call: GPU parity tests, bench cfg2 (+cfg3), ncu launch list, one full ncu capture of the counting kernel.
# usage: tools/gpu_round2.sh TAG [cfg3=0|1] [ncu=0|1]
TAG=${1:-run}; CFG3=${2:-1}; NCU=${3:-1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.err
cat gpurun_out/${TAG}_bench_cfg2.json
if [ "$CFG3" = "1" ]; then
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 600 gpurun_out/${TAG}_bench_cfg3.err
cat gpurun_out/${TAG}_bench_cfg3.json
fi
if [ "$NCU" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qs_count_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_rows100 python tools/profile_count.py --n 100 --m 10000 > gpurun_out/${TAG}_rows100.log 2>&1
fi
ls -la gpurun_out/ | tail -8
