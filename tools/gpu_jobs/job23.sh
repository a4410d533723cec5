timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t13.log
tail -5 gpurun_out/t13.log
