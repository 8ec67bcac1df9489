#!/bin/bash
# C5 on the full-resolution synthetic grid (1987x2449x1351, HBM-resident): launch times of both layouts, lookup counters, and the DRAM bytes
# of one launch of each layout from a single-pass ncu capture (dram__bytes only: a multi-pass capture would save / restore ~80 GB per pass)
mkdir -p gpurun_out
# K19 with the texel coordinate folded into one fma per axis (SKY_K19_FOLD=1) and the collision probability in one multiplication (=2)
export SPP=8,64 DIGEST=1 GRID_SCALE=1
python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_fold_r02H.log
for v in fold1 fold2; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/pt_timing.py 2>&1 | tee -a gpurun_out/k19_fold_r02H.log; done
(time SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_fold2.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "path_tracer or c5_path") > gpurun_out/pt_tests_fold2_r02H.log 2>&1; echo "fold2 tests rc=$?"; tail -5 gpurun_out/pt_tests_fold2_r02H.log
export GRID_SCALE=16 SPP=4 TAG=r02H DIGEST=
(time timeout 900 python tools/c5_large.py) > gpurun_out/c5_full_r02H.log 2>&1; echo "timing rc=$?"; tail -30 gpurun_out/c5_full_r02H.log
for hw in 0 1; do
  (time MODE=launch HW=$hw timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k19_path_trace -c 1 --csv \
     --log-file gpurun_out/c5_full_hw${hw}_r02H.csv python tools/c5_large.py) > gpurun_out/c5_full_ncu_hw${hw}_r02H.log 2>&1
  echo "ncu hw=$hw rc=$?"; tail -4 gpurun_out/c5_full_hw${hw}_r02H.csv
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
