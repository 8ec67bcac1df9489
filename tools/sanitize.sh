#!/bin/bash
# compute-sanitizer over tools/sanitize_smoke.py; summaries into gpurun_out/ (copied to profiles/ by hand).   usage: tools/sanitize.sh TAG
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
    (time timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py) > gpurun_out/sanitize_${tool}_$tag.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize: done" gpurun_out/sanitize_${tool}_$tag.log
done
