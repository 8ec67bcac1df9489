#!/bin/bash
# final single-GPU pass of the round: the whole GPU suite, smoke(), the default bench line
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -q) > gpurun_out/gputests_r02_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/gputests_r02_final.log
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_r02_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02_final.log
(time timeout 900 python bench.py) > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02_final_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_final_n1.json"))
f = d["frame_4k"]
print("pt", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "frame", f["ms_per_frame"], f["parts_ms"], "k16 frac", f["roofline"]["frac"], "K6", f["roofline_K6"]["frac"], "obj", f["object_shading_variant"]["ms_per_frame"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a == "frame_ms" or a == "parts_us" or a == "gsamples_per_s"})
PY
