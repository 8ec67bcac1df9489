#!/bin/bash
mkdir -p gpurun_out
PARITY=1 timeout 300 python tools/k6_ab.py 2>&1 | tee gpurun_out/k6_froxeltex_r02w.log
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ibl.py tests/test_gpu_earth.py -m gpu -q -x -k "composite or c2 or c3 or star or object or frame_with or pcss or pipelin or overlap or ragged") > gpurun_out/gputests_r02w.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gputests_r02w.log
