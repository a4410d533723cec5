#!/bin/bash
# round 2, job 14: k_burg heads + carried denominator with the direct-sum redo as a separate loop; td-iir e2e with larger chunks; sweeps with the
# conditioning probe and the contraction-free fp64 scan
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "burg or vad or full_length or fwss or afterFB or sweep" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest14.log
for v in "1 4" "1 3"; do
  set -- $v
  CTU_BURG_REC=$1 CTU_BURG_MINB=$2 python bench.py --workload fwss_burg --others none --steps 5 --no-cpu-baseline --e2e-steps 0 --cli-utts 0 > gpurun_out/r2_burg_h2_$1_$2.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2_burg_h2_$1_$2.json')); print('rec $1 minb $2', d['ms_per_step'], d['kernel_ms_per_step'], d.get('selfcheck'))"
done
python bench.py --workload tdiir --others none --steps 5 --no-cpu-baseline --e2e-steps 2 --cli-utts 0 > gpurun_out/r2_bench_tdiir.json 2> gpurun_out/r2_bench_tdiir.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_tdiir.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tdiir.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
for spec in "200 23" "200 61 0.3 1.0" "200 63 0.5 0.7"; do
  timeout 1200 python tools/parity_sweep.py gpu $spec > "gpurun_out/r2c_sweep_$(echo $spec | tr ' ' '_').txt" 2>&1; echo "sweep $spec rc=$?"; tail -1 "gpurun_out/r2c_sweep_$(echo $spec | tr ' ' '_').txt"
  grep "^   -fs" "gpurun_out/r2c_sweep_$(echo $spec | tr ' ' '_').txt" | grep -v hwss
done
