timeout 1500 python tools/parity_sweep.py gpu 180 31 2>/dev/null | grep -v "^$" > gpurun_out/sweep_gpu3.log
grep -A1 "CUDA" gpurun_out/sweep_gpu3.log | grep -v "^--" | paste - - | grep -v hwss | cut -c1-900; tail -1 gpurun_out/sweep_gpu3.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "sweep" 2>&1 | tail -3
