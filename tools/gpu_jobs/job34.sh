timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/t17.log
tail -4 gpurun_out/t17.log
python tools/cli_e2e.py 4000 2>&1 | tail -6
