#!/bin/bash
# after sky_launch_count (host-side counter only): a quick GPU subset, smoke, and the default bench line with the counted gpu_launches
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q -x -k "launch_count or error_behaviour or stream_parity or lut_bake_parity or cloud_chain_parity") > gpurun_out/gputests_r02R.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests_r02R.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_r02R.log 2>&1; echo "smoke rc=$?"
(time timeout 600 python bench.py) > gpurun_out/bench_r02R_n1.json 2> gpurun_out/bench_r02R_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02R_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02R_n1.json").read().strip().splitlines()[-1])
print("pt", d["value"], "e2e", d["e2e"]["value"], "gpu_launches", d["gpu_launches"], "frame", d["frame_4k"]["ms_per_frame"], "frame launches", d["frame_4k"]["gpu_launches"], "c5_full", d["configs"]["c5_full"].get("gsamples_per_s"))
PY
