B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000"
tools/gpu_jobs/ncu_cap.sh q_spec "k_frames2" 3 1 $B --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh q_fea "k_frames" 3 1 $B --workload mfcc_d_a
rm -f gpurun_out/src_q_*.csv
