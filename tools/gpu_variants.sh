#!/bin/bash
# time the counting kernel of every tuning build in tools/variants/ on cfg2 and larger workloads
for lib in tools/variants/*.so; do
  echo "== $lib"
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --chunks 0 --ring 32x8 2>&1 | tail -1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0 --ring 32x8 2>&1 | tail -1
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 500 --m 1000 --seed 3000 --p-missing 0.1 --p-contract 0.05 --chunks 0 --ring 32x8 2>&1 | tail -1
done
