# contract check + ncu evidence for the 8f.3 workload (stacking behind the MFCC front end)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000"
python bench.py --workload mfcc_trap5 --no-cpu-baseline > gpurun_out/b_trap5.json 2> gpurun_out/b_trap5.err; tail -2 gpurun_out/b_trap5.err
$B --workload mfcc_trap5 > gpurun_out/plain_mfcc_trap5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_mfcc_trap5.csv $B --workload mfcc_trap5 > gpurun_out/nl_mfcc_trap5.log 2>&1
tools/gpu_jobs/ncu_cap.sh p_mfcc_trap5 "k_stack" 3 1 $B --workload mfcc_trap5
rm -f gpurun_out/srccu_p_*.csv gpurun_out/src_p_*.csv
du -sh gpurun_out
