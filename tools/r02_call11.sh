#!/bin/bash
# full GPU suite (K7 ground pass new), the FFMA2 issue probe, one bench line
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/probes/ffma2_probe.cu && /tmp/ffma2_probe > gpurun_out/ffma2_probe.log 2>&1; cat gpurun_out/ffma2_probe.log
(time timeout 1800 python -m pytest tests -m gpu -q -x) > gpurun_out/gputests_r02k.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gputests_r02k.log
