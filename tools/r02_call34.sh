#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/lut_coop_variants_r02E.log
: > $out
timeout 300 python tools/lut_coop_timing.py >> $out 2>&1
for v in occ1024 occ1024u2 occ768u2 occ512u2; do
  SKYB200_LIB=skyrendering_b200/csrc/variant_$v.so timeout 300 python tools/lut_coop_timing.py >> $out 2>&1
done
cat $out
# per-kernel durations of the default build (lanes 8)
SCENES=c3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_lut_r02E.csv python tools/lut_coop_timing.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_lut_r02E.csv | tee gpurun_out/launches_lut_r02E.md
