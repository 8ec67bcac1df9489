#!/bin/bash
# cooperative LUT march: timing + error against the oracle, and the exact LUT tests
mkdir -p gpurun_out
timeout 900 python tools/lut_coop_probe.py > gpurun_out/lut_coop_r02D.log 2>&1; echo "probe rc=$?"
cat gpurun_out/lut_coop_r02D.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lut_bake_parity or reference_shader_digests" 2>&1 | tail -3
