#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_earth.py tests/test_gpu_ibl.py -m gpu -q -x) > gpurun_out/gputests_r02z.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gputests_r02z.log
(time timeout 900 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline) > gpurun_out/bench_r02z_n1.json 2> gpurun_out/bench_r02z_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02z_n1.json"))
f = d["frame_4k"]
print("pt", d["value"], "frame", f["ms_per_frame"], f["parts_ms"], "k16 frac", f["roofline"]["frac"], "K6", f["roofline_K6"]["frac"], f["roofline_K17_K18"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a == "frame_ms" or a == "parts_us"})
PY
