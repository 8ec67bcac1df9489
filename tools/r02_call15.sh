#!/bin/bash
mkdir -p gpurun_out
SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_k16stats.so timeout 120 python tools/k16_stats.py 2>&1 | tee gpurun_out/k16_stats_r02m.log
