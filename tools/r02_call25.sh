#!/bin/bash
mkdir -p gpurun_out
HW=1 timeout 600 tools/ncu_one.sh k17 k17_reconstruct 6 r02v -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k18 k18_upscale 6 r02v -- python tools/ncu_frame.py
SCENE=c2 timeout 600 tools/ncu_one.sh k6c2 k6_composite 6 r02v -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k13 k13_shadow_froxel 6 r02v -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k11 k11_shadow_map 6 r02v -- python tools/ncu_frame.py
ls gpurun_out | grep r02v
