#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/k16_persist_r02F.log
: > $out
EXACT_LUTS=1 timeout 120 python tools/frame_prod.py >> $out 2>&1
for prio in 0 1; do for p in 0 2 3 4 5 6; do
  SKYB200_K16_PERSIST=$p SKYB200_LANE2_PRIORITY=$prio timeout 120 python tools/frame_prod.py >> $out 2>&1
done; done
cat $out
