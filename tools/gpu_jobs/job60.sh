B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh burg3 "k_burg" 3 1 $B --utts 500 --workload fwss_burg
rm -f gpurun_out/src_burg3.csv
