#!/bin/bash
# ncu evidence for the IBL chain and the object branch of K6 (summaries only; see tools/ncu_round.sh).  usage: tools/ncu_ibl.sh TAG
tag=$1
out=gpurun_out
mkdir -p $out
cap() {  # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/${name}_$tag -f "$@" > $out/ncu_${name}_$tag.log 2>&1
    python tools/ncu_summary.py /tmp/${name}_$tag.ncu-rep > $out/${name}_$tag.md 2>> $out/ncu_${name}_$tag.log
    ncu -i /tmp/${name}_$tag.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${name}_dump.csv 2>> $out/ncu_${name}_$tag.log
    python tools/ncu_hot_lines.py /tmp/${name}_dump.csv 30 > $out/${name}_${tag}_hotlines.md 2>> $out/ncu_${name}_$tag.log
}
export OBJECTS=1 HW=1
cap k24 k24_prefilter 4 python tools/ncu_frame.py
cap k6obj k6_composite 4 python tools/ncu_frame.py
cap k23 k23_env_sh 4 python tools/ncu_frame.py
ls -la $out | tail -8
