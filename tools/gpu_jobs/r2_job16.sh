#!/bin/bash
# round 2, job 16: ncu evidence of the final kernels: launch lists (gpu__time_duration) and one --set full capture per workload
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck --cli-utts 0"
for w in mfcc_exten mfcc_d_a plp trapdct exten fwss_burg tdiir; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 || { echo "plain failed $w"; tail -3 gpurun_out/plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_r02_$w.csv $B --workload $w > gpurun_out/nl_$w.log 2>&1
done
tools/gpu_jobs/ncu_cap.sh p_mfcc_exten "k_frames|k_bank|k_delta" 12 3 $B --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh p_fwss_burg "k_burg|k_cepdet" 8 2 $B --utts 500 --workload fwss_burg
tools/gpu_jobs/ncu_cap.sh p_exten "k_frames|k_nr_scan|k_synth" 12 3 $B --workload exten
tools/gpu_jobs/ncu_cap.sh p_tdiir "k_tdiir" 8 2 $B --workload tdiir
tools/gpu_jobs/ncu_cap.sh p_plp "k_bank|k_lpc" 8 2 $B --workload plp
tools/gpu_jobs/ncu_cap.sh p_trapdct "k_bank|k_trapdct" 8 2 $B --workload trapdct
rm -f gpurun_out/srccu_p_mfcc_exten.csv gpurun_out/srccu_p_exten.csv gpurun_out/srccu_p_plp.csv gpurun_out/srccu_p_trapdct.csv gpurun_out/src_p_plp.csv gpurun_out/src_p_trapdct.csv
du -sh gpurun_out
