#!/bin/bash
# round 2, job 35: the fp64 general kernels with the padded complex layout (bank conflicts): parity at the other sampling rates, timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "8k or 11k or 22k or 32k or 44k or burg or raw or wave or sweep or afterFB or lpa or lpc or cepdist" > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest35.log
T="python tools/time_args.py 4000 --"
( echo "== 8 kHz exten raw"; $T -fs 8000 -format_in raw -preset exten -format_out raw
  echo "== 8 kHz fwss burg"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode fwss -vad burg -format_out htk
  echo "== 8 kHz mfcc exten afterFB (fp64 band path)"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode exten -nr_when afterFB -format_out htk
  echo "== 44.1 kHz exten wave"; $T -fs 44100 -format_in raw -preset exten -format_out raw ) > gpurun_out/r2_any_times9.txt 2>&1
grep -v "host path" gpurun_out/r2_any_times9.txt
