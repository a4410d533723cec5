#!/bin/bash
# round 2, job 32: what the general fp64 kernels (k_burg_any, k_synth_frames_any) wait for: ncu on 8 kHz runs
mkdir -p gpurun_out
T="python tools/time_args.py 1000 --"
tools/gpu_jobs/ncu_cap.sh p_burg_any "k_burg_any" 2 1 $T -fs 8000 -format_in raw -preset mfcc -preem 0.97 -nr_mode fwss -vad burg -format_out htk
tools/gpu_jobs/ncu_cap.sh p_synth_any "k_synth_frames_any" 2 1 $T -fs 8000 -format_in raw -preset exten -format_out raw
rm -f gpurun_out/srccu_p_synth_any.csv
python profiles/summarize_ncu.py gpurun_out/raw_p_burg_any.csv
python profiles/summarize_srccu.py gpurun_out/srccu_p_burg_any.csv | head -40
python profiles/summarize_ncu.py gpurun_out/raw_p_synth_any.csv
