#pragma once
#include "common.cuh"

int dense_gemv(const double* A, const double* x, double* y, int64_t m, int64_t n, cudaStream_t st);
// LU row-major (M x M) with L unit-lower / U upper as LAPACK getrf leaves them; perm[i] = source row of P b
int dense_lu_solve(const double* LU, const int* perm, const double* b, double* x, int M, cudaStream_t st);
int csr_spmv_d(const int* indptr, const int* indices, const double* data, const double* x, double* y, int64_t nrows, cudaStream_t st);
int csr_spmv_z(const int* indptr, const int* indices, const cplx* data, const cplx* x, cplx* y, int64_t nrows, cudaStream_t st);
int vec_axpby(double a, const double* x, double b, const double* y, double* out, int64_t n, cudaStream_t st);
int vec_mul(const double* x, const double* y, double* out, int64_t n, cudaStream_t st);
int rowvec_apply(const double* row, const double* fr, double* out, int M, int64_t n, cudaStream_t st);
int scalar_jumps(const double* bv, const double* bx, const double* by, const double* nx, const double* ny, const double* ur,
                 const double* in_est, int M, int64_t n, double sign, double* tau, cudaStream_t st);
// structured QFS apply (qfs_structured.py): banded + low-block complex mat-vec in Fourier space
int banded_apply_z(const cplx* xh, int n_in, int64_t N, const cplx* band, const cplx* lowT, int n_out, int BH, int KL,
                   int64_t Ks, cplx* yh, int64_t M, cudaStream_t st);
int real_to_complex(const double* x, cplx* z, int64_t n, cudaStream_t st);
int complex_parts(const cplx* z, double* re, double* im, int64_t n, cudaStream_t st);
int complex_parts_strided(const cplx* z, double* re, double* im, int64_t n, int64_t stride, int accumulate, cudaStream_t st);
