#!/bin/bash
# L2 fetch granularity (sky_set_l2_fetch_granularity) on the grids beyond L2: launch times of both layouts at 32 / 64 / 128 bytes for the
# full-resolution and the quarter grid, and the DRAM bytes of one launch at 32 bytes (single-pass ncu, dram__bytes only)
mkdir -p gpurun_out
export TAG=r02I
(time GRID_SCALE=16 SPP=4 GRANULARITY=128,64,32 FINAL_GRANULARITY=32 timeout 900 python tools/c5_large.py) > gpurun_out/c5_full_r02I.log 2>&1; echo "x16 rc=$?"; grep -E "granularity|grid " gpurun_out/c5_full_r02I.log
(time GRID_SCALE=4 SPP=8 GRANULARITY=128,64,32 FINAL_GRANULARITY=32 timeout 600 python tools/c5_large.py) > gpurun_out/c5_quarter_r02I.log 2>&1; echo "x4 rc=$?"; grep -E "granularity|grid " gpurun_out/c5_quarter_r02I.log
for hw in 0 1; do
  (time GRID_SCALE=16 SPP=4 MODE=launch HW=$hw FINAL_GRANULARITY=32 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k19_path_trace -c 1 --csv \
     --log-file gpurun_out/c5_full_g32_hw${hw}_r02I.csv python tools/c5_large.py) > gpurun_out/c5_full_ncu_g32_hw${hw}_r02I.log 2>&1
  echo "ncu hw=$hw rc=$?"; tail -3 gpurun_out/c5_full_g32_hw${hw}_r02I.csv | cut -d, -f5,13-15
done
# the sixteenth-size (L2-resident) bench workload must not care; streaming kernels of the 4K frame neither
SPP=64 GRID_SCALE=1 python tools/pt_timing.py 2>&1 | tee gpurun_out/k19_default_r02I.log
