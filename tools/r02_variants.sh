#!/bin/bash
# times K14-K16 for every skyrendering_b200/csrc/variant_*.so named in $@ (and the default library first)
mkdir -p gpurun_out
log=gpurun_out/k16_variants_${TAG:-r02}.log
python tools/k16_ab.py 2>&1 | tee $log
for v in "$@"; do
  SKYB200_LIB=$PWD/skyrendering_b200/csrc/variant_$v.so python tools/k16_ab.py 2>&1 | tee -a $log
done
