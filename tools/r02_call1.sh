#!/bin/bash
# Round 2, first GPU call: the whole GPU suite (incl. the new BASELINE-size parity tests), one default bench line, and
# fresh ncu --set full summaries (with stall reasons) of the two kernels the round is about.
mkdir -p gpurun_out
nvidia-smi -L; nproc
(time timeout 1800 python -m pytest tests -m gpu -q --durations=15) > gpurun_out/gputests_r02a.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/gputests_r02a.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/bench_r02a_n1.json 2> gpurun_out/bench_r02a_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02a_n1.err
HW=1 tools/ncu_one.sh k16hw k16_render 6 r02a -- python tools/ncu_frame.py
tools/ncu_one.sh k6 k6_composite 6 r02a -- python tools/ncu_frame.py
python tools/ncu_traffic.py k6_composite=/tmp/k6_r02a.ncu-rep k16_render_hw=/tmp/k16hw_r02a.ncu-rep > /dev/null 2>&1
cp profiles/traffic_r02.json gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -20
