#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wave_shapes or cloud_chain or strict_arithmetic or full_size") > gpurun_out/gputests_r02s.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/gputests_r02s.log
timeout 200 python tools/k16_ab.py 2>&1 | tee gpurun_out/k16_r02s.log
