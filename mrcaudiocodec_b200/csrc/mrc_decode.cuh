// mrc_decode.cuh -- declarations of the decode kernels (defined in mrc_decode.cu).
#pragma once
#include "mrc_internal.cuh"
