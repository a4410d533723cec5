// Translation unit of the standalone noise-reduction scans, the VAD module and the synthesis (ctu_nr_kernels.cuh).
// This unit keeps the SCALAR complex arithmetic (ctu_fft.cuh): with the packed FP32 instructions the synthesis kernel measured
// 11.0 ms (three CTAs per SM, 24-36 bytes spilled under its 80-register cap) and 15.5 ms (two CTAs, 111 registers, no
// spills) against 8.65 ms scalar per 6.24 M frames (tools/gpu_jobs/r2_job42.sh, r2_job43.sh) -- register pairs must be
// aligned, and the inverse split keeps 2 x 8 bin pairs plus the 16-point column live at once.
#define CTU_NO_F32X2 1
#include "ctu_nr_kernels.cuh"
