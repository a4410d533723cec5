#!/bin/bash
# round 2, job 18: td-iir after the staging / dependency changes; where the general (non-512-point) kernels stand: 8 kHz MFCC / PLP / exten -> waveform, 11 and 44 kHz
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tdiir or td_iir" > gpurun_out/r2_pytest18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest18.log
python bench.py --workload tdiir --others none --steps 5 --no-cpu-baseline --e2e-steps 2 --cli-utts 0 > gpurun_out/r2_bench_tdiir.json 2> gpurun_out/r2_bench_tdiir.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_tdiir.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tdiir.json')); print(d['value'], d['ms_per_step'], d['selfcheck'], d['selfcheck_detail'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
T="python tools/time_args.py 4000 --"
( echo "== 8 kHz mfcc d_a"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
  echo "== 8 kHz plp"; $T -fs 8000 -format_in raw -preset plpc -format_out htk
  echo "== 8 kHz mfcc exten"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode exten -fea_delta d_a -format_out htk
  echo "== 8 kHz exten raw"; $T -fs 8000 -format_in raw -preset exten -format_out raw
  echo "== 8 kHz fwss burg"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode fwss -vad burg -format_out htk
  echo "== 44.1 kHz mfcc"; $T -fs 44100 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
  echo "== 16 kHz mfcc d_a (reference point)"; $T -fs 16000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk ) > gpurun_out/r2_any_times.txt 2>&1
cat gpurun_out/r2_any_times.txt
