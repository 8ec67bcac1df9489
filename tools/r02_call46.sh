#!/bin/bash
# 8-GPU (and 4-GPU) lines of the final build
mkdir -p gpurun_out
for n in 8 4; do
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3) > gpurun_out/bench_r02P_n$n.json 2> gpurun_out/bench_r02P_n$n.err; echo "bench n$n rc=$?"; tail -2 gpurun_out/bench_r02P_n$n.err
  python - $n <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_r02P_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "frame_4k_ms", "sharded_equals_single")})
PY
done
