timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t22.log
tail -5 gpurun_out/t22.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_main.json 2>gpurun_out/b_main.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload fwss_burg --utts 4000 > gpurun_out/b_burg.json 2>gpurun_out/b_burg.err
