#pragma once
#include "nufft.cuh"
// cmat: (M, M) DCT matrix (device); fr: (M, n) device; coef_work: (M, n) scratch; theta/t: targets;
// out[out_pos ? out_pos[p] : p] = value
int radial_to_points(Nufft1Plan* plan, const double* cmat, int M, const double* fr, double* coef_work,
                     const double* theta, const double* t, int64_t npts, double* out, const int64_t* out_pos,
                     cudaStream_t st);
