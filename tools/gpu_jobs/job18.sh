timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t11.log
tail -3 gpurun_out/t11.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_mfcc.json 2>gpurun_out/b_mfcc.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload plp > gpurun_out/b_plp.json 2>gpurun_out/b_plp.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh f1 "k_frames" 3 1 $B --utts 2000 --workload mfcc_d_a
