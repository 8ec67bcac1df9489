#!/bin/bash
# K19: the PCG hashes of the hot block with their constant shifts as IMAD.HI (ALU pipe -> FMA pipe)
mkdir -p gpurun_out
export SPP=8,64 DIGEST=1 GRID_SCALE=1
python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_hash_r02O.log
for v in himad; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/pt_timing.py 2>&1 | tee -a gpurun_out/k19_hash_r02O.log; done
