#!/bin/bash
# One round's ncu evidence, summarised ON THE BOX (the .ncu-rep files stay in /tmp: three of them exceed gpurun's 64 MiB return limit).
#   tools/ncu_round.sh TAG   ->   gpurun_out/{launches_TAG.csv, k16_TAG.md, k16hw_TAG.md, k19_TAG.md, *_hotlines.md, traffic_TAG.json}
tag=$1
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --spp 8 --skip-cpu-baseline > $out/bench_under_ncu_$tag.log 2>&1
cap() {  # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/${name}_$tag -f "$@" > $out/ncu_${name}_$tag.log 2>&1
    python tools/ncu_summary.py /tmp/${name}_$tag.ncu-rep > $out/${name}_$tag.md 2>> $out/ncu_${name}_$tag.log
    ncu -i /tmp/${name}_$tag.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${name}_dump.csv 2>> $out/ncu_${name}_$tag.log
    python tools/ncu_hot_lines.py /tmp/${name}_dump.csv 30 > $out/${name}_${tag}_hotlines.md 2>> $out/ncu_${name}_$tag.log
}
cap k16 k16_render 6 python tools/ncu_frame.py
HW=1 cap k16hw k16_render 6 python tools/ncu_frame.py
cap k6 k6_composite 6 python tools/ncu_frame.py
cap k2 k2_multiscattering 6 python tools/ncu_frame.py
cap k4 k4_aerial 6 python tools/ncu_frame.py
SPP=${K19_SPP:-16} cap k19 k19_path_trace 1 python tools/pt_timing.py
ls -la $out | tail -20
