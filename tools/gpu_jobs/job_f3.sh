set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f3_pytest.log 2>&1; tail -5 gpurun_out/f3_pytest.log
timeout 300 python tools/time_fea_chain.py > gpurun_out/f3_time.log 2>&1; tail -12 gpurun_out/f3_time.log
