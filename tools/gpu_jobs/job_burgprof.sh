B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 500"
tools/gpu_jobs/ncu_cap.sh p_burg "k_burg" 3 1 $B --workload fwss_burg
python profiles/summarize_srccu.py gpurun_out/srccu_p_burg.csv > gpurun_out/burg_src.txt 2>&1
rm -f gpurun_out/srccu_p_*.csv gpurun_out/src_p_*.csv
