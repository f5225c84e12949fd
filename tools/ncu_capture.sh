#!/usr/bin/env bash
# usage (on the GPU box): tools/ncu_capture.sh <tag> <kernel regex> <skip> <count> [source]
# Captures `ncu --set full` of tools/ncu_target.py and leaves small CSV pages in gpurun_out/ (the .ncu-rep is
# kept only when it is < 12 MB: gpurun_out/ is capped at 64 MiB).
set -u
tag=$1; rx=$2; skip=$3; cnt=$4; src=${5:-}
extra=""; [ -n "$src" ] && extra="--import-source on"
ncu --set full --clock-control none $extra -k "regex:$rx" -s "$skip" -c "$cnt" -f -o "gpurun_out/$tag" \
    python tools/ncu_target.py "${NCU_BATCH:-64}" > "gpurun_out/$tag.log" 2>&1
ncu -i "gpurun_out/$tag.ncu-rep" --page raw --csv > "gpurun_out/${tag}_raw.csv" 2>/dev/null
[ -n "$src" ] && ncu -i "gpurun_out/$tag.ncu-rep" --page source --csv > "gpurun_out/${tag}_src.csv" 2>/dev/null
sz=$(stat -c %s "gpurun_out/$tag.ncu-rep" 2>/dev/null || echo 0)
[ "$sz" -gt 12000000 ] && rm -f "gpurun_out/$tag.ncu-rep"
tail -n 2 "gpurun_out/$tag.log"
