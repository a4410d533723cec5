timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t21.log
tail -12 gpurun_out/t21.log
