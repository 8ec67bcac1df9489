#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3) > gpurun_out/bench_r02q_n$N.json 2> gpurun_out/bench_r02q_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02q_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02q_n$N.json").read().strip().splitlines()[-1])
print("pt", d["value"], "frame_4k_ms", d["frame_4k_ms"], "sharded_equals_single", d["sharded_equals_single"], d["sharded_checks"])
print(d["frame_4k"]["parts_ms"])
PY
