CTU_SPEC_GEN1=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_main1.json 2>gpurun_out/b_main1.err
CTU_SPEC_GEN1=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload exten > gpurun_out/b_exten1.json 2>gpurun_out/b_exten1.err
