timeout 1200 python tools/parity_sweep.py gpu 200 23 2>/dev/null | grep -v "^$" > gpurun_out/sweep_gpu2.log
grep -A1 "CUDA" gpurun_out/sweep_gpu2.log | grep -v hwss | grep -B1 "htk" | tail -20; tail -1 gpurun_out/sweep_gpu2.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
