#!/bin/bash
mkdir -p gpurun_out
SIZES=3840x2160,1920x1080 python tools/k16_ab.py 2>&1 | tee gpurun_out/k16_wave_r02h.log
SIZES=3840x2160,1920x1080 SKYB200_K16_LITERAL=1 python tools/k16_ab.py 2>&1 | tee -a gpurun_out/k16_wave_r02h.log
(time timeout 1800 python -m pytest tests -m gpu -q) > gpurun_out/gputests_r02h.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/gputests_r02h.log
