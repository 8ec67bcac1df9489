#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
(time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x) > gpurun_out/gputests_multi_r02q.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gputests_multi_r02q.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3) > gpurun_out/bench_r02q_n2.json 2> gpurun_out/bench_r02q_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r02q_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02q_n2.json").read().strip().splitlines()[-1])
print("pt", d["value"], "frame_4k_ms", d["frame_4k_ms"], "sharded_equals_single", d["sharded_equals_single"], d["sharded_checks"])
PY
