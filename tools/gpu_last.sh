#!/bin/bash
TAG=${1:-last}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 300 gpurun_out/${TAG}_bench_cfg2.err; cat gpurun_out/${TAG}_bench_cfg2.json
timeout 900 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; cat gpurun_out/${TAG}_bench_cfg3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qs_count_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_rows100 python tools/profile_count.py --n 100 --m 10000 > gpurun_out/${TAG}_rows100.log 2>&1
