B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh f3x "k_frames3" 6 2 $B --utts 2000 --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh f3 "k_frames3" 3 1 $B --utts 2000 --workload mfcc_d_a
du -sh gpurun_out
