#!/bin/bash
# ncu --set full of K19 after the fold / branchless changes (16 kFrameIds: the launch shape only matters for the drain share), with source hot lines
mkdir -p gpurun_out
SPP=16 GRID_SCALE=1 timeout 900 tools/ncu_one.sh k19 k19_path_trace 1 r02M -- python tools/pt_timing.py
cat gpurun_out/k19_r02M.md | head -30
