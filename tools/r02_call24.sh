#!/bin/bash
# DRAM traffic of K19 on the 497x612x338 grid: corner-packed cells (exact) vs the R8 3-D array through the texture unit
mkdir -p gpurun_out
for hw in 0 1; do
  tag=c5q_hw$hw
  HW=$hw GRID_SCALE=4 SPP=8 timeout 900 ncu --set full --clock-control none -k regex:k19_path_trace -s 1 -c 1 -o /tmp/$tag -f python tools/pt_timing.py > gpurun_out/ncu_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/k19_${tag}_r02u.md 2>> gpurun_out/ncu_$tag.log
done
TRAFFIC_FILE=traffic_r02.json python tools/ncu_traffic.py k19_path_trace_c5_quarter=/tmp/c5q_hw0.ncu-rep k19_path_trace_c5_quarter_tex=/tmp/c5q_hw1.ncu-rep | tail -22
cp profiles/traffic_r02.json gpurun_out/traffic_r02.json
for hw in 0 1; do HW=$hw GRID_SCALE=4 SPP=8 timeout 300 python tools/pt_timing.py 2>&1 | tail -1; done
