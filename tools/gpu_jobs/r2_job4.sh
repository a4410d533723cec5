#!/bin/bash
# round 2, job 4: chunked fused scan; ncu captures of the current kernels; the three sweep seeds with open non-hwss classes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest4.log
python bench.py --no-cpu-baseline --e2e-steps 1 --others none > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench4.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step'])"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck"
tools/gpu_jobs/ncu_cap.sh r2_mfcc_exten "k_frames|k_bank|k_delta" 12 3 $B --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh r2_fwss_burg "k_burg|k_cepdet" 8 2 $B --utts 500 --workload fwss_burg
tools/gpu_jobs/ncu_cap.sh r2_exten "k_frames|k_nr_scan|k_synth" 12 3 $B --workload exten
rm -f gpurun_out/srccu_r2_mfcc_exten.csv gpurun_out/srccu_r2_exten.csv
for spec in "200 23" "200 61 0.3 1.0" "200 63 0.5 0.7"; do
  timeout 900 python tools/parity_sweep.py gpu $spec > "gpurun_out/r2_sweep_$(echo $spec | tr ' ' '_').txt" 2>&1; echo "sweep $spec rc=$?"; tail -2 "gpurun_out/r2_sweep_$(echo $spec | tr ' ' '_').txt"
done
du -sh gpurun_out
