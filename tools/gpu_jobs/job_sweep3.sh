timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "11k or 8k" > gpurun_out/odd_pytest.log 2>&1; tail -5 gpurun_out/odd_pytest.log
timeout 1500 python tools/parity_sweep.py gpu 200 61 0.3 1.0 > gpurun_out/sweep_gpu_fs61.log 2>&1; tail -2 gpurun_out/sweep_gpu_fs61.log
timeout 1500 python tools/parity_sweep.py gpu 200 63 0.5 0.7 > gpurun_out/sweep_gpu_fs63.log 2>&1; tail -2 gpurun_out/sweep_gpu_fs63.log
