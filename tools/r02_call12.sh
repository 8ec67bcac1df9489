#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/probes/ffma2_probe.cu && /tmp/ffma2_probe > gpurun_out/ffma2_probe.log 2>&1; cat gpurun_out/ffma2_probe.log
(time python tools/sanitize_smoke.py) 2>&1 | tail -12
tools/sanitize.sh r02
