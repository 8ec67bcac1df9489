#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/ncu_one.sh k6 k6_composite 6 r02B -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k16hw k16_render 6 r02B -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k17 k17_reconstruct 6 r02B -- python tools/ncu_frame.py
TRAFFIC_FILE=traffic_r02.json python tools/ncu_traffic.py k6_composite=/tmp/k6_r02B.ncu-rep k16_render_hw=/tmp/k16hw_r02B.ncu-rep > /dev/null
cp profiles/traffic_r02.json gpurun_out/traffic_r02.json
grep -E "duration|warp instructions|issue slots|eligible" gpurun_out/k6_r02B.md gpurun_out/k16hw_r02B.md gpurun_out/k17_r02B.md
