#!/bin/bash
for lib in quartetscores_b200/libqscuda.so tools/variants/libqs_noxr.so tools/variants/libqs_xrfirst.so tools/variants/libqs_runf.so; do
  echo "== $lib"
  QS_DEBUG_PLAN=1 QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --chunks 0,3,7 --ring 32x8 2>&1 | grep -v "^\[qscuda\]" ; QS_DEBUG_PLAN=1 QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --chunks 0 --ring 32x8 2>&1 | grep "^\[qscuda\]" | sort | uniq -c | head -4
  QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0,2 --ring 32x8 2>&1
  QS_DEBUG_PLAN=1 QSCUDA_LIB=$PWD/$lib python tools/sweep_chunks.py --n 200 --m 3000 --seed 2100 --chunks 0 --ring 32x8 2>&1 | grep "^\[qscuda\]" | sort | uniq -c | head -4
done > gpurun_out/r02_i_variants.txt 2>&1
cat gpurun_out/r02_i_variants.txt
