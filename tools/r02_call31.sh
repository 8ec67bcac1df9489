#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k6_variants_r02A.log
timeout 200 python tools/k6_ab.py 2>&1 | tee $log
for v in rcp3 rcp3o4 o4 u1 u2 u1o4 u2o4; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so timeout 200 python tools/k6_ab.py 2>&1 | grep c3 | tee -a $log; done
