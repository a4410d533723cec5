#!/bin/bash
# round 2, job 1: parity suite after the ADVICE fixes / k_frames2 changes, default bench line, copy ceiling on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r2_pytest1.log
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tools/copy_probe.sh 1 gpurun_out/r2_copy_probe_n1.jsonl 3 > /dev/null 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'self',d['selfcheck'],'e2e',d['e2e'])
print(d['kernel_ms_per_step'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('selfcheck'), v.get('kernel_ms_per_step'), v.get('error'))
PY
