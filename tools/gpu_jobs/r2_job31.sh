#!/bin/bash
# round 2, job 31: the general Burg kernel with the lattice in registers (windows up to 512 samples)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "burg or vad or sweep or carry or 8k or 11k or 44k" > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest31.log
T="python tools/time_args.py 4000 --"
( echo "== 8 kHz fwss burg"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode fwss -vad burg -format_out htk
  echo "== 11 kHz fwss burg 32/16"; $T -fs 11025 -format_in raw -preset mfcc -preem 0.97 -w 32 -s 16 -nr_mode fwss -vad burg -format_out htk
  echo "== 8 kHz vad cepdist lpc"; $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -format_out htk -vad_out_mode vad -vad_thr_mode adapt -vad_cri_mode cepdist -vad_cepdist_mode lpc -vad burg ) > gpurun_out/r2_any_times8.txt 2>&1
cat gpurun_out/r2_any_times8.txt
