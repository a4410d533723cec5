#!/bin/bash
# round 2, job 47: validation of the packed-arithmetic build (the driver's round-end sequence, r2_job26.sh), then the ncu evidence of
# the kernels that changed: launch lists and one --set full capture for mfcc_exten, trapdct, tdiir
bash tools/gpu_jobs/r2_job26.sh
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck --cli-utts 0"
for w in mfcc_exten trapdct tdiir; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 || { echo "plain failed $w"; tail -3 gpurun_out/plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_r02_$w.csv $B --workload $w > gpurun_out/nl_$w.log 2>&1
done
tools/gpu_jobs/ncu_cap.sh p_mfcc_exten "k_frames|k_bank|k_delta" 12 3 $B --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh p_trapdct "k_bank|k_trapdct" 8 2 $B --workload trapdct
tools/gpu_jobs/ncu_cap.sh p_tdiir_full "k_tdiir" 2 2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0 --workload tdiir
rm -f gpurun_out/srccu_p_mfcc_exten.csv gpurun_out/srccu_p_trapdct.csv gpurun_out/src_p_trapdct.csv gpurun_out/srccu_p_tdiir_full.csv
du -sh gpurun_out
