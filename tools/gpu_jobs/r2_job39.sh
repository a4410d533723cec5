#!/bin/bash
# round 2, job 39: ncu summaries of the kernels changed after job 16 (k_burg with the Newton division, k_lpc, k_frames256 + the general spec -> fea kernel at 8 kHz)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck --cli-utts 0"
tools/gpu_jobs/ncu_cap.sh p_fwss_burg "k_burg|k_cepdet" 8 2 $B --utts 500 --workload fwss_burg
tools/gpu_jobs/ncu_cap.sh p_plp "k_bank|k_lpc" 8 2 $B --workload plp
tools/gpu_jobs/ncu_cap.sh p_mfcc8k "k_frames256|k_frames_any|k_delta" 6 3 python tools/time_args.py 2000 -- -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
rm -f gpurun_out/srccu_p_plp.csv gpurun_out/src_p_plp.csv gpurun_out/srccu_p_mfcc8k.csv
for w in fwss_burg plp; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_r02_$w.csv $B --workload $w > gpurun_out/nl_$w.log 2>&1
done
du -sh gpurun_out
