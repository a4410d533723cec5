// Translation unit of the standalone noise-reduction scans, the VAD module and the synthesis (ctu_nr_kernels.cuh).
#include "ctu_nr_kernels.cuh"
