timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload fwss_burg > gpurun_out/b_burg.json 2>gpurun_out/b_burg.err
