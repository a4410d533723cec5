timeout 1500 python tools/parity_sweep.py gpu 250 71 0.3 0.3 > gpurun_out/sweep_gpu_71.log 2>&1; tail -1 gpurun_out/sweep_gpu_71.log
timeout 1500 python tools/parity_sweep.py gpu 250 72 0.0 0.0 > gpurun_out/sweep_gpu_72.log 2>&1; tail -1 gpurun_out/sweep_gpu_72.log
