# one compute-sanitizer tool per call (B200_PROFILING.md); smallest cases that touch every kernel family
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "mfcc_exten_d_a or plpc_ark or exten_raw or fwss_burg_pfile or trapdct_51_8 or vad_cepdist_lpc_adapt or mfcc_E_d_a or ragged or vad_perc_d_a_drop" > gpurun_out/memcheck.log 2>&1
echo "exit $?" >> gpurun_out/memcheck.log
tail -15 gpurun_out/memcheck.log
