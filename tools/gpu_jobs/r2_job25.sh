#!/bin/bash
# round 2, job 25: k_tdiir_filter at bench scale under ncu (occupancy, pipes, stalls)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0 --workload tdiir"
tools/gpu_jobs/ncu_cap.sh p_tdiir_full "k_tdiir_filter" 3 1 $B
rm -f gpurun_out/srccu_p_tdiir_full.csv
python profiles/summarize_ncu.py gpurun_out/raw_p_tdiir_full.csv
python profiles/summarize_src.py gpurun_out/src_p_tdiir_full.csv | head -40
