#pragma once
#include "common.cuh"

int layer_set_k0_table(const double* cheb0, const double* cheb1, int ncheb);
int layer_eval_k0(const double* x, double* out, int64_t n, int order, cudaStream_t st);
// kind 0 Laplace, 1 modified Helmholtz (parameter k); q = charge * weight
int layer_apply_scalar(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                       const double* tx, const double* ty, int64_t ntrg, double* out, int accumulate, cudaStream_t st);
// normal derivative of the same sums at targets with unit normals (tnx, tny)
int layer_apply_scalar_dn(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                          const double* tx, const double* ty, const double* tnx, const double* tny, int64_t ntrg, double* out,
                          int accumulate, cudaStream_t st);
// out_uvp: (3, ntrg)
int layer_apply_stokes(const double* sx, const double* sy, const double* fx, const double* fy, int64_t nsrc,
                       const double* tx, const double* ty, int64_t ntrg, double* out_uvp, int accumulate, cudaStream_t st);
int layer_kernel_matrix(int kind, double k, int layer, const double* tx, const double* ty, int64_t nt, const double* sx,
                        const double* sy, const double* snx, const double* sny, const double* w, int64_t ns, double* out,
                        cudaStream_t st);
