#!/usr/bin/env bash
# usage: ncu_cap.sh <name> <kernel-regex> <skip> <count> <cmd...>
# Plain run first (must exit 0), then one `ncu --set full` capture; the report is exported
# to CSV pages (raw metrics + per-source-line stalls) and removed, so gpurun_out stays small.
name=$1; kre=$2; skip=$3; cnt=$4; shift 4
"$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run failed for $name"; tail -5 gpurun_out/plain_$name.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/raw_$name.csv 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page source --csv > gpurun_out/src_$name.csv 2>/dev/null
ncu -i /tmp/prof_$name.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/srccu_$name.csv 2>/dev/null
ls -la /tmp/prof_$name.ncu-rep gpurun_out/raw_$name.csv gpurun_out/src_$name.csv
rm -f /tmp/prof_$name.ncu-rep
