#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/k16_group_r02r.log
for g in 8 4; do
  SKYB200_K16_GROUP=$g SIZES=1920x1080,3840x2160 timeout 300 python tools/k16_ab.py 2>&1 | sed "s/^/group=$g /" | tee -a $log
done
(time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cloud_chain or strict_arithmetic or full_size or c3 or c4 or overlap or pipelin or hw_filter or ragged or viewport") > gpurun_out/gputests_r02r.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gputests_r02r.log
