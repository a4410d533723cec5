#!/bin/bash
# round 2, job 8: full parity suite with device-formatted container rows, VAD debug files, final k_burg variant; bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest8.log
python bench.py --no-cpu-baseline --e2e-steps 1 --others none --workload fwss_burg --steps 5 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench8.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step'])"
