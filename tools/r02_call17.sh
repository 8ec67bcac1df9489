#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/ncu_one.sh k6 k6_composite 6 r02o -- python tools/ncu_frame.py
HW=1 timeout 600 tools/ncu_one.sh k16hw k16_render 6 r02o -- python tools/ncu_frame.py
tail -3 gpurun_out/ncu_k6_r02o.log
