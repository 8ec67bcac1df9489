#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k19_variants_r02y.log
DIGEST=1 SPP=8,64 timeout 200 python tools/pt_timing.py 2>&1 | tee $log
for v in mt8 mt16 mt24 tr4 tr16 tr4mt16; do DIGEST=1 SPP=8,64 SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so timeout 200 python tools/pt_timing.py 2>&1 | tee -a $log; done
