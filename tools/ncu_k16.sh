#!/bin/bash
# ncu --set full captures (one launch each): K16 exact + hardware-filtered (4K, scene c3), K19 (720p, 2 spp).  tools/ncu_k16.sh TAG
tag=$1
ncu --set full --clock-control none --import-source on -k regex:k16_render -s 6 -c 1 -o gpurun_out/k16_$tag -f python tools/ncu_frame.py > gpurun_out/ncu_k16_$tag.log 2>&1
HW=1 ncu --set full --clock-control none --import-source on -k regex:k16_render -s 6 -c 1 -o gpurun_out/k16hw_$tag -f python tools/ncu_frame.py > gpurun_out/ncu_k16hw_$tag.log 2>&1
# K19 in the launch shape of bench.py's roofline: 1280x720 x 64 kFrameIds (second launch = after the warm-up call)
[ -n "$SKIP_K19" ] || SPP=64 ncu --set full --clock-control none --import-source on -k regex:k19_path_trace -s 1 -c 1 -o gpurun_out/k19_$tag -f python tools/pt_timing.py > gpurun_out/ncu_k19_$tag.log 2>&1
