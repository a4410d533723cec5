#!/bin/bash
# round 2, job 38: the 256-point front end after its FFT pieces moved to ctu_fft.cuh (CPU-emulated there): 8 kHz parity + timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "8k or 8khz or g711 or sweep" > gpurun_out/r2_pytest40.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest38.log
python tools/time_args.py 4000 -- -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
