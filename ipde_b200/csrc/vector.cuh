#pragma once
#include "fft.cuh"

int vec_uv_rt(const double* fu, const double* fv, const double* nx, const double* ny, const double* tx, const double* ty, int M,
              int n, double* out0, double* out1, int inverse, cudaStream_t st);
int vec_radial_traction(const double* Ur, const double* Ut, const double* dUr, const double* P, const double* D00,
                        const double* est, const double* rsp, const double* nx, const double* ny, const double* tx,
                        const double* ty, int M, int n, double* Tx, double* Ty, cudaStream_t st);
int vec_jumps(const double* rtx, const double* rty, const double* txx, const double* txy, const double* tyy, const double* bu,
              const double* bv, const double* inx, const double* iny, int n, double sign, double* tau, cudaStream_t st);
int vec_theta_derivative(FftPlan* p, const double* f, double* work_cplx, double* out, int rows, cudaStream_t st);
int vec_pressure_continuity(const double* txx, const double* tyy, const double* pb, const double* est, double* pr, int M, int n,
                            cudaStream_t st);
