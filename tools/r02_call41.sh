#!/bin/bash
# K19 retune after the branchless batch: collisions per batch, blocks per SM, rounds per visit
mkdir -p gpurun_out
export SPP=64 DIGEST=1 GRID_SCALE=1
python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_retune_r02L.log
for v in b6 b6o5 occ7 occ5 tr16 tr4; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/pt_timing.py 2>&1 | tee -a gpurun_out/k19_retune_r02L.log; done
