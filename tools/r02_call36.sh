#!/bin/bash
# round-2 verification pass after the cooperative LUT march: GPU suite, smoke, sanitizer, ncu captures (K19 at the bench's launch shape for
# the issue roof; the two cooperative LUT kernels), default bench line
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/gputests_r02G.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputests_r02G.log
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_r02G.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02G.log
tools/sanitize.sh r02G
SPP=64 timeout 900 tools/ncu_one.sh k19 k19_path_trace 1 r02G -- python tools/pt_timing.py
SCENES=c3 timeout 300 tools/ncu_one.sh k34coop k34_cooperative 3 r02G -- python tools/lut_coop_timing.py
SCENES=c3 timeout 300 tools/ncu_one.sh k2coop k2_multiscattering_cooperative 3 r02G -- python tools/lut_coop_timing.py
TRAFFIC_FILE=traffic_r02.json python tools/ncu_traffic.py k19_path_trace=/tmp/k19_r02G.ncu-rep > /dev/null
cp profiles/traffic_r02.json gpurun_out/traffic_r02.json
grep -E "duration|warp instructions|issue slots|eligible|registers" gpurun_out/k19_r02G.md gpurun_out/k34coop_r02G.md gpurun_out/k2coop_r02G.md
(time timeout 1200 python bench.py) > gpurun_out/bench_r02G_n1.json 2> gpurun_out/bench_r02G_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02G_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02G_n1.json").read().strip().splitlines()[-1])
f = d["frame_4k"]
print("pt", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "issue", d["roofline"].get("issue"))
print("frame", f["ms_per_frame"], f["parts_ms"], "k16 frac", f["roofline"]["frac"], "K6", f["roofline_K6"]["frac"], "lutvar", f["lut_arithmetic_variant"], "obj", f["object_shading_variant"]["ms_per_frame"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a in ("frame_ms", "parts_us", "gsamples_per_s", "cooperative")})
PY
