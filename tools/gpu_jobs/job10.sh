timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t7.log
timeout 800 python bench.py --steps 5 --warmup 3 --others --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -5 gpurun_out/t7.log
