timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t16.log
tail -3 gpurun_out/t16.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload exten > gpurun_out/b_exten.json 2>gpurun_out/b_exten.err
