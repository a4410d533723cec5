python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_main.json 2>gpurun_out/b_main.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload exten > gpurun_out/b_exten.json 2>gpurun_out/b_exten.err
