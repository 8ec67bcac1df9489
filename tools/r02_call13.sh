#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k6_variants_r02l.log
PARITY=1 python tools/k6_ab.py 2>&1 | tee $log
for v in k6u1 k6u2 k6u4 k6u1o4 k6u2o4; do
  SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/k6_ab.py 2>&1 | tee -a $log
done
