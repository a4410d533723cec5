// Translation unit of the fp64 frame kernels (ctu_precise.cuh): band-domain noise reduction, ill-conditioned LPC,
// features that feed VAD decisions.
#define CTU_PRECISE_IMPL
#include "ctu_precise.cuh"
