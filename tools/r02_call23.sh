#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/lut_variants_r02t.log
timeout 200 python tools/lut_ab.py 2>&1 | tee $log
for v in m1 m4; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so timeout 200 python tools/lut_ab.py 2>&1 | tee -a $log; done
