#!/bin/bash
mkdir -p gpurun_out
SIZES=3840x2160,1920x1080 python tools/k16_ab.py 2>&1 | tee gpurun_out/k16_wave_r02i.log
for v in r32 la8 o5 o6 b64; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/k16_ab.py 2>&1 | tee -a gpurun_out/k16_wave_r02i.log; done
PARITY=1 SIZES=1920x1080 python tools/k16_ab.py 2>&1 | grep render | tee -a gpurun_out/k16_wave_r02i.log
HW=1 tools/ncu_one.sh k16hw k16_render 6 r02i -- python tools/ncu_frame.py
