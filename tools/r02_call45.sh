#!/bin/bash
# 2-GPU sanity of the default bench line after the K19 changes: sharded path tracer + sharded 4K frame, sharded_equals_single
mkdir -p gpurun_out
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/bench_r02P_n2.json 2> gpurun_out/bench_r02P_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_r02P_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02P_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "frame_4k_ms", "sharded_equals_single", "sharded_checks", "scaling")})
PY
(time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x) > gpurun_out/gpumulti_r02P.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/gpumulti_r02P.log
