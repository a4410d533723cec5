timeout 1200 python tools/parity_sweep.py gpu 200 21 > gpurun_out/sweep_gpu_21.log 2>&1; tail -3 gpurun_out/sweep_gpu_21.log
timeout 1200 python tools/parity_sweep.py gpu 200 22 > gpurun_out/sweep_gpu_22.log 2>&1; tail -3 gpurun_out/sweep_gpu_22.log
