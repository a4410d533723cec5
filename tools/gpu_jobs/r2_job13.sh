#!/bin/bash
# round 2, job 13: td-iir-mfcc parity + timing; k_burg with per-thread heads (A/B: carried denominator on / off, 4 or 3 CTAs per SM)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tdiir or td_iir or burg or vad or full_length or fwss" > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest13.log
python bench.py --workload tdiir --others none --steps 5 --no-cpu-baseline --e2e-steps 1 --cli-utts 0 > gpurun_out/r2_bench_tdiir.json 2> gpurun_out/r2_bench_tdiir.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_tdiir.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tdiir.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d.get('selfcheck_detail'), d['kernel_ms_per_step'], d['e2e'])"
for v in "1 4" "0 4" "1 3"; do
  set -- $v
  CTU_BURG_REC=$1 CTU_BURG_MINB=$2 python bench.py --workload fwss_burg --others none --steps 5 --no-cpu-baseline --e2e-steps 0 --cli-utts 0 > gpurun_out/r2_burg_h_$1_$2.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2_burg_h_$1_$2.json')); print('rec $1 minb $2', d['ms_per_step'], d['kernel_ms_per_step'], d.get('selfcheck'))"
done
