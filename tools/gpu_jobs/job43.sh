timeout 1200 python tools/parity_sweep.py gpu 160 11 2>/dev/null | grep -v "^$" > gpurun_out/sweep_gpu.log
tail -40 gpurun_out/sweep_gpu.log
