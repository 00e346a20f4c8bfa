#!/bin/bash
# last validation of the committed state: the new distance-kernel test, the whole GPU suite, smoke, both default bench lines
TAG=${1:-r02_last}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "warp_distance" > gpurun_out/${TAG}_newtest.log 2>&1; tail -3 gpurun_out/${TAG}_newtest.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_cfg2_1gpu.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 300 gpurun_out/${TAG}_bench_cfg2.err; cut -c1-260 gpurun_out/${TAG}_bench_cfg2_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg3_1gpu.json 2> gpurun_out/${TAG}_bench_cfg3.err; cut -c1-260 gpurun_out/${TAG}_bench_cfg3_1gpu.json
