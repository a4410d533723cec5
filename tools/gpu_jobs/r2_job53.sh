#!/bin/bash
# round 2, job 53: the bit-identity test of the compact / in-place cepstra layouts
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "compact_and_in_place" > gpurun_out/r2_pytest53.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest53.log
