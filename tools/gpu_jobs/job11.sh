B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
tools/gpu_jobs/ncu_cap.sh frames2 "k_frames2" 3 1 $B --utts 2000 --workload mfcc_d_a
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_mfcc.json 2>gpurun_out/b_mfcc.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload plp > gpurun_out/b_plp.json 2>gpurun_out/b_plp.err
du -sh gpurun_out
