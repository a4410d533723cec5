// Translation unit of the time-domain IIR filter-bank features (ctu_tdiir.cuh).
#define CTU_TDIIR_IMPL
#include "ctu_tdiir.cuh"
