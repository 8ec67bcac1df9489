#!/bin/bash
# K19 after the coordinate fold: resident blocks per SM (4 / 5 / 6; 6 fits in 80 registers without spilling now) and 3 collisions per batch
mkdir -p gpurun_out
export SPP=8,64 DIGEST=1 GRID_SCALE=1
python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_occ_r02J.log
for v in occ6 occ4 b3; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/pt_timing.py 2>&1 | tee -a gpurun_out/k19_occ_r02J.log; done
