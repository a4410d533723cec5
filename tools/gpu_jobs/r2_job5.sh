#!/bin/bash
# round 2, job 5: k_burg (zeroing by switch, persistent), k_bank (approximate reciprocals in the scan, branch-free chunks), tightened parity tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest5.log
python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench5.json'))
print('value',d['value'],'ms',d['ms_per_step'],'self',d['selfcheck'], d.get('selfcheck_detail'))
print(d['kernel_ms_per_step'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('selfcheck'), v.get('kernel_ms_per_step'), v.get('error'), (v.get('selfcheck_detail') or {}).get('oracle_max_err_over_tol'))
PY
tail -3 gpurun_out/r2_bench5.err
