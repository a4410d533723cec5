#!/bin/bash
# round 2, job 52: the default bench line of the final build and the ncu capture of its kernels (compact cepstra, k_bank without the
# end-of-tile barrier)
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench52.json 2> gpurun_out/r2_bench52.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench52.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck --cli-utts 0"
tools/gpu_jobs/ncu_cap.sh p_mfcc_exten "k_frames|k_bank|k_delta" 12 3 $B --workload mfcc_exten
rm -f gpurun_out/srccu_p_mfcc_exten.csv
