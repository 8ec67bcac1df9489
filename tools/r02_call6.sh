#!/bin/bash
mkdir -p gpurun_out
PARITY=1 python tools/k16_ab.py 2>&1 | tee gpurun_out/k16_wave_r02g.log
for v in la4 la4o8 o6 o8 la2o8; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/k16_ab.py 2>&1 | tee -a gpurun_out/k16_wave_r02g.log; done
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -k "cloud_chain or ragged or voxel_realtime or hardware_filtering or overlap or pipelining or banded or host_buffer or c3_cloud or c4_cloud or full_size") > gpurun_out/gputests_r02g.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gputests_r02g.log
HW=1 tools/ncu_one.sh k16hw k16_render 6 r02g -- python tools/ncu_frame.py
