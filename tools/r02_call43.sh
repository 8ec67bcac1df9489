#!/bin/bash
# round-2 verification pass after the K19 fold / branchless work, the JPEG reader and the full-grid bench entry: GPU suite, smoke, a fresh
# ncu capture of K19 at the bench's launch shape (issue roof + traffic), the default bench line
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/gputests_r02N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputests_r02N.log
(time timeout 600 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_r02N.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r02N.log
SPP=64 timeout 900 tools/ncu_one.sh k19 k19_path_trace 1 r02N -- python tools/pt_timing.py
TRAFFIC_FILE=traffic_r02.json python tools/ncu_traffic.py k19_path_trace=/tmp/k19_r02N.ncu-rep > /dev/null
cp profiles/traffic_r02.json gpurun_out/traffic_r02.json
grep -E "duration|warp instructions|issue slots|eligible|registers|active lanes" gpurun_out/k19_r02N.md
(time timeout 1200 python bench.py) > gpurun_out/bench_r02N_n1.json 2> gpurun_out/bench_r02N_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02N_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02N_n1.json").read().strip().splitlines()[-1])
f = d["frame_4k"]
print("pt", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "issue", d["roofline"].get("issue"))
print("frame", f["ms_per_frame"], f["parts_ms"], "k16 frac", f["roofline"]["frac"], "K6", f["roofline_K6"]["frac"])
for k, v in d["configs"].items():
    print(k, {a: b for a, b in v.items() if a.endswith("_us") or a.endswith("_ms") or a in ("frame_ms", "parts_us", "gsamples_per_s", "skipped", "ms_per_launch")}, v.get("roofline", {}).get("frac") if isinstance(v.get("roofline"), dict) else "")
PY
(time timeout 600 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/bench_r02N_ref.json 2> gpurun_out/bench_r02N_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_r02N_ref.json
