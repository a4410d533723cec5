#!/bin/bash
# round 2, job 6: k_burg variant A/B (clearing: masks / switch / select chain x persistent or not), VAD debug side files
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vad or burg" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest6.log
for v in 0 1 2 10 11 12; do
  CTU_BURG_VARIANT=$v python bench.py --workload fwss_burg --others none --utts 4000 --steps 5 --no-cpu-baseline --e2e-steps 0 --no-selfcheck > gpurun_out/r2_burg_v$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2_burg_v$v.json')); print('variant $v', d['ms_per_step'], d['kernel_ms_per_step'].get('k_burg'))"
done
