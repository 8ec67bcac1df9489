#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02z.csv python bench.py --steps 1 --warmup 1 --spp 8 --skip-cpu-baseline > gpurun_out/bench_under_ncu_r02z.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02z.csv > gpurun_out/launches_r02z.md 2>&1; head -60 gpurun_out/launches_r02z.md
