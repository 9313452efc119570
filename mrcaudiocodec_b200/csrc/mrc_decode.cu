// mrc_decode.cu -- decode kernels (K5).  Filled in below.
#include "mrc_decode.cuh"
