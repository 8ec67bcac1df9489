#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x -k "cloud_chain or ragged or voxel_realtime or hardware_filtering or overlap or pipelining or banded or host_buffer or c3_cloud or c4_cloud or full_size") > gpurun_out/gputests_r02c.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gputests_r02c.log
PARITY=1 python tools/k16_ab.py 2>&1 | tee gpurun_out/k16_ab_r02c.log
SKYB200_K16_LITERAL=1 python tools/k16_ab.py 2>&1 | tee -a gpurun_out/k16_ab_r02c.log
HW=1 tools/ncu_one.sh k16hw k16_render 6 r02c -- python tools/ncu_frame.py
