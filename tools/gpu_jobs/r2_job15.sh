#!/bin/bash
# round 2, job 15: full GPU suite (with the 200-set sweeps, td-iir-mfcc, the contraction-free fp64 scan), default bench line, td-iir timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest15.log
python bench.py > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench15.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench15.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step']); print(d.get('cli')); print(d['cpu_baseline'])
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('kernel_ms_per_step'), v.get('selfcheck'), (v.get('e2e') or {}).get('value'), v.get('error'))"
