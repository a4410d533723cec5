// Translation unit of the Burg cepstral detector (ctu_burg.cuh).
#include "ctu_burg.cuh"
