timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/t23.log
tail -3 gpurun_out/t23.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others > gpurun_out/b_all.json 2>gpurun_out/b_all.err
