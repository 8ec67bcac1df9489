#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k6_variants_r02m.log
PARITY=1 python tools/k6_ab.py 2>&1 | tee $log
for v in k6l5 k6l8; do
  SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/k6_ab.py 2>&1 | tee -a $log
done
SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_k16stats.so python tools/k16_stats.py 2>&1 | tee gpurun_out/k16_stats_r02m.log
