timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t8.log
tail -3 gpurun_out/t8.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh burg2 "k_burg" 3 1 $B --utts 500 --workload fwss_burg
tools/gpu_jobs/ncu_cap.sh synth2 "k_synth" 3 1 $B --utts 1000 --workload exten
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_mfcc.json 2>gpurun_out/b_mfcc.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload trapdct > gpurun_out/b_trap.json 2>gpurun_out/b_trap.err
du -sh gpurun_out
