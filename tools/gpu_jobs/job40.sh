timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t19.log
tail -3 gpurun_out/t19.log
