#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for lib in tools/variants/libqs_sb*.so; do echo "== $lib"; QSCUDA_LIB=$PWD/$lib python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 2>&1 | tail -1; done > gpurun_out/${TAG}_score_variants.txt 2>&1
cat gpurun_out/${TAG}_score_variants.txt
