#!/bin/bash
# round 2, job 21 (run again as job 22 with the transposed DCT matrix in shared memory and radix-4 passes): general frame kernel parity + timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "8k or 11k or 22k or 32k or 44k or dc1 or dither or sweep or g711 or burg or raw or wave" > gpurun_out/r2_pytest29.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest29.log
T="python tools/time_args.py 4000 --"
( echo "== 8 kHz mfcc d_a"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
  echo "== 8 kHz plp"; $T -fs 8000 -format_in raw -preset plpc -format_out htk
  echo "== 8 kHz mfcc exten"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode exten -fea_delta d_a -format_out htk
  echo "== 44.1 kHz mfcc"; $T -fs 44100 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk
  echo "== 22.05 kHz mfcc"; $T -fs 22050 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk ) > gpurun_out/r2_any_times7.txt 2>&1
grep -v "^   k_delta\|^   k_lpc" gpurun_out/r2_any_times7.txt
T="python tools/time_args.py 4000 --"
( echo "== 8 kHz exten raw"; $T -fs 8000 -format_in raw -preset exten -format_out raw
  echo "== 8 kHz fwss burg"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode fwss -vad burg -format_out htk ) > gpurun_out/r2_any_times7b.txt 2>&1
cat gpurun_out/r2_any_times7b.txt
