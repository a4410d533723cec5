timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "burg or vad or sweep" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --workload fwss_burg --utts 4000 > gpurun_out/b_burg.json 2>gpurun_out/b_burg.err
