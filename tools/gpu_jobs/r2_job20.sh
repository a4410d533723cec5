#!/bin/bash
# round 2, job 20: k_burg with the Newton division and reciprocal constants in the cepstrum recursion
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "burg or vad or full_length or fwss or carry or sweep" > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest20.log
python bench.py --workload fwss_burg --others none --steps 5 --no-cpu-baseline --e2e-steps 1 --cli-utts 0 > gpurun_out/r2_burg20.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_burg20.json')); print(d['ms_per_step'], d['kernel_ms_per_step'], d.get('selfcheck'), d.get('selfcheck_detail'))"
