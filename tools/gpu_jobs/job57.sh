timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/t24.log
tail -3 gpurun_out/t24.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload fwss_burg --utts 4000 > gpurun_out/b_burg.json 2>gpurun_out/b_burg.err
