#!/bin/bash
# K19: resolution of a batch as selects (SKY_K19_BRANCHLESS=1) and the blends too (=2); digests must equal the default's
mkdir -p gpurun_out
export SPP=8,64 DIGEST=1 GRID_SCALE=1
python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_branchless_r02K.log
for v in bl1 bl2; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/pt_timing.py 2>&1 | tee -a gpurun_out/k19_branchless_r02K.log; done
