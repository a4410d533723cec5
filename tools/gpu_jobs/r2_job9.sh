#!/bin/bash
# round 2, job 9: k_cepdet with the reference cepstrum in registers; default bench line (all workloads + the CLI file -> file block)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "burg or vad or cepdist or full_length" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest9.log
python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench9.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench9.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step']); print(d.get('cli')); print(d['cpu_baseline'])
for k,v in d['workloads'].items(): print(k, v['value'], v['ms_per_step'], v['kernel_ms_per_step'])"
