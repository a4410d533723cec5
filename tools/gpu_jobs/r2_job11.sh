#!/bin/bash
# round 2, job 11: k_burg with the carried denominator (A/B against the direct sums), decisions bit-exact; the three 200-set sweeps again
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vad or burg or full_length or fwss" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest11.log
for v in 1 0; do
  CTU_BURG_REC=$v python bench.py --workload fwss_burg --others none --steps 5 --no-cpu-baseline --e2e-steps 0 --cli-utts 0 > gpurun_out/r2_burg_rec$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2_burg_rec$v.json')); print('rec $v', d['ms_per_step'], d['kernel_ms_per_step'], d.get('selfcheck'))"
done
for spec in "200 23" "200 61 0.3 1.0" "200 63 0.5 0.7"; do
  timeout 900 python tools/parity_sweep.py gpu $spec > "gpurun_out/r2b_sweep_$(echo $spec | tr ' ' '_').txt" 2>&1; echo "sweep $spec rc=$?"; tail -1 "gpurun_out/r2b_sweep_$(echo $spec | tr ' ' '_').txt"
  grep "^   -fs" "gpurun_out/r2b_sweep_$(echo $spec | tr ' ' '_').txt" | grep -v hwss
done
