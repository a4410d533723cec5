#!/bin/bash
# round 2, job 26 (and again as job 33 at the end of the round): the driver's round-end sequence: full GPU suite, smoke, reference arm, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest51.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest51.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref51.json 2> gpurun_out/r2_ref51.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_ref51.json
python bench.py > gpurun_out/r2_bench51.json 2> gpurun_out/r2_bench51.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench51.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench51.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step'], d['gpu_launches']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['frac_of_copy_only']); print(d.get('cli')); print(d['cpu_baseline']); print(d['clocks'])
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('selfcheck'), (v.get('e2e') or {}).get('value'), v.get('error'))"
