#!/bin/bash
# round 2, job 30: compute-sanitizer memcheck over the kernels added or reshaped in round 2 (k_bank incl. the fused scan, k_nr_scan4,
# k_tdiir_*, k_frames256, k_nr_scan_carry, k_frames_any with staged tables, row formatting, VAD debug, k_burg)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
  -k "test_cuda_matches_reference_golden and (mfcc_exten_d_a or plpc_ark or exten_raw or fwss_burg_pfile or trapdct_51_8 or tdiir_w25s10 or tdiir_8k or mfcc8k_d_a or plp8k or fwss_burg_8k or exten_raw_8k or logspec32k or mfcc44k or vaddbg_perc or mfcc_dither1 or mfcc_dc1 or hwss_burg_spec_pow) or carry_fwss_burg or carry_fwss_file_afterFB or carry_fwss_raw or 8khz_front" \
  > gpurun_out/r2_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r2_memcheck.log
tail -12 gpurun_out/r2_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_cli.py -m gpu -q -x -p no:cacheprovider -k "pfile or htk_files or td_iir" > gpurun_out/r2_memcheck_cli.log 2>&1
echo "exit $?" >> gpurun_out/r2_memcheck_cli.log
tail -5 gpurun_out/r2_memcheck_cli.log
