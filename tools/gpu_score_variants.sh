#!/bin/bash
for lib in tools/variants/libqs_*.so; do echo "== $lib"; QSCUDA_LIB=$PWD/$lib python tools/profile_count.py --n 500 --m 300 --seed 3000 --p-missing 0.1 --p-contract 0.05 --reps 3 2>&1 | tail -1; done
