set -x
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/t5.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh synth k_synth 3 1 $B --utts 1000 --workload exten
tools/gpu_jobs/ncu_cap.sh burg k_burg 3 1 $B --utts 500 --workload fwss_burg
tools/gpu_jobs/ncu_cap.sh frames "k_frames|k_delta" 6 2 $B --utts 2000 --workload mfcc_d_a
tools/gpu_jobs/ncu_cap.sh trap k_trapdct 3 1 $B --utts 2000 --workload trapdct
tail -3 gpurun_out/t5.log
du -sh gpurun_out
