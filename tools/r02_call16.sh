#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k16_variants_r02n.log
PARITY=1 timeout 300 python tools/k16_ab.py 2>&1 | tee $log
SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_eager.so timeout 200 python tools/k16_ab.py 2>&1 | tee -a $log
SIZES=1920x1080 timeout 200 python tools/k16_ab.py 2>&1 | tee -a $log
