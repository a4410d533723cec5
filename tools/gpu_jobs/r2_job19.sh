#!/bin/bash
# round 2, job 19: Burg detector at 25 ms as two kernels (transform | lattice with one frame per warp) against the single kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "burg or vad or full_length or fwss or carry" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest19.log
for v in 1 0; do
  CTU_BURG_SPLIT=$v python bench.py --workload fwss_burg --others none --steps 5 --no-cpu-baseline --e2e-steps 1 --cli-utts 0 > gpurun_out/r2_burg_split$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2_burg_split$v.json')); print('split $v', d['ms_per_step'], d['kernel_ms_per_step'], d.get('selfcheck'), d['e2e']['ms_per_step'])"
done
