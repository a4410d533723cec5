#!/bin/bash
# round 2, job 54: ncu capture of the final td-iir filter kernel (warp-private staging) at bench scale, one launch
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k "regex:k_tdiir_filter" -s 1 -c 1 -f -o /tmp/prof_td python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0 --workload tdiir > gpurun_out/ncu_p_tdiir_final.log 2>&1
ncu -i /tmp/prof_td.ncu-rep --page raw --csv > gpurun_out/raw_p_tdiir_final.csv 2>/dev/null
ncu -i /tmp/prof_td.ncu-rep --page source --csv > gpurun_out/src_p_tdiir_final.csv 2>/dev/null
ls -la gpurun_out/raw_p_tdiir_final.csv gpurun_out/src_p_tdiir_final.csv
