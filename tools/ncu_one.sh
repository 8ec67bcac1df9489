#!/bin/bash
# tools/ncu_one.sh NAME KERNEL_REGEX SKIP TAG -- command...   (summaries only; the .ncu-rep stays in /tmp)
name=$1 regex=$2 skip=$3 tag=$4; shift 5
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/${name}_$tag -f "$@" > $out/ncu_${name}_$tag.log 2>&1
python tools/ncu_summary.py /tmp/${name}_$tag.ncu-rep > $out/${name}_$tag.md 2>> $out/ncu_${name}_$tag.log
ncu -i /tmp/${name}_$tag.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${name}_dump.csv 2>> $out/ncu_${name}_$tag.log
python tools/ncu_hot_lines.py /tmp/${name}_dump.csv 30 > $out/${name}_${tag}_hotlines.md 2>> $out/ncu_${name}_$tag.log
