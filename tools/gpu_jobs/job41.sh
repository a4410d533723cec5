timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t20.log
tail -6 gpurun_out/t20.log
