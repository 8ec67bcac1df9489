#!/bin/bash
# K6 fast march: parity of everything that touches the composite, its timing, an ncu capture; and the LUT-only instantiation (scene c2)
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ibl.py -m gpu -q -k "c2 or c3 or c4 or c5_path or composite or star or optional or strict_arithmetic_frames or object or pcss or cloud_chain or cpp_frame") > gpurun_out/gputests_r02b.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/gputests_r02b.log
(time timeout 900 python bench.py --steps 1 --warmup 1 --spp 16 --skip-cpu-baseline) > gpurun_out/bench_r02b_n1.json 2> gpurun_out/bench_r02b_n1.err; echo "bench rc=$?"
tools/ncu_one.sh k6 k6_composite 6 r02b -- python tools/ncu_frame.py
SCENE=c2 tools/ncu_one.sh k6c2 k6_composite 6 r02b -- python tools/ncu_frame.py
python tools/ncu_traffic.py k6_composite=/tmp/k6_r02b.ncu-rep > /dev/null 2>&1
cp profiles/traffic_r02.json gpurun_out/ 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02b_n1.json"))
f = d["frame_4k"]
print("frame", f["ms_per_frame"], f["parts_ms"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a == "frame_ms"})
PY
