#!/bin/bash
mkdir -p gpurun_out
PARITY=1 timeout 300 python tools/k6_ab.py 2>&1 | tee gpurun_out/k6_halfluts_r02x.log
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/gputests_r02x.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gputests_r02x.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/bench_r02x_n1.json 2> gpurun_out/bench_r02x_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02x_n1.json"))
f = d["frame_4k"]
print("pt", d["value"], "frame", f["ms_per_frame"], f["parts_ms"], "k16 frac", f["roofline"]["frac"], "K6", f["roofline_K6"]["frac"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a == "frame_ms" or a == "parts_us"})
PY
