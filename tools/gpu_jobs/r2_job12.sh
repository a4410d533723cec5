#!/bin/bash
# round 2, job 12: td-iir-mfcc parity + timing; where the two open sweep classes differ
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tdiir or td_iir" > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest12.log
python bench.py --workload tdiir --others none --steps 5 --no-cpu-baseline --e2e-steps 1 --cli-utts 0 > gpurun_out/r2_bench_tdiir.json 2> gpurun_out/r2_bench_tdiir.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_tdiir.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tdiir.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d.get('selfcheck_detail'), d['kernel_ms_per_step'], d['e2e'])"
python bench.py --impl reference --workload tdiir --steps 2 --warmup 1 > gpurun_out/r2_ref_tdiir.json 2> gpurun_out/r2_ref_tdiir.err; echo "ref rc=$?"; cat gpurun_out/r2_ref_tdiir.json | cut -c1-400
python tools/diag_sweep_miss.py > gpurun_out/r2_diag_sweep.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/r2_diag_sweep.txt | grep -v Warning | head -60
