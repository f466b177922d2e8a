#pragma once
#include "fft.cuh"
#include "gmres.cuh"

struct AnnularPlan;

// All matrices row-major doubles on the host; kinv_mode_minor[(j*M + k)*n + i] = Kinv_i[j][k].
int annular_plan_create(int n, int M, double k, const double* D01, const double* D12, const double* R01, const double* R12,
                        const double* R02, const double* ibc, const double* obc, const double* ks, const double* psi1,
                        const double* ipsi1, const double* ipsi2, const double* kinv_mode_minor, AnnularPlan** out);
void annular_plan_destroy(AnnularPlan* p);
int annular_apply(AnnularPlan* p, const cplx* in, cplx* out, cudaStream_t st);
int annular_precond(AnnularPlan* p, const cplx* in, cplx* out, cudaStream_t st);
// f: (M, n) real device; ig/og: (n) real device or NULL (zero); f_sign = -1 for Poisson (annular/poisson.py:20)
int annular_solve(AnnularPlan* p, const double* f, const double* ig, const double* og, double f_sign, double tol,
                  int restart, int maxiter, double* u_out, int* iters, double* relres, cudaStream_t st);

// ---- annular Stokes (annular_stokes.cu).  Host arrays; kinv_mode_minor[((j*R + k)*n + i)] complex, R = 3M-1,
// zero at the Nyquist mode i = n/2.  ks[n/2] must be 0.
struct AnnularStokesPlan;
int annular_stokes_plan_create(int n, int M, double mu, const double* D01, const double* D12, const double* R01,
                               const double* R12, const double* R02, const double* P10, const double* ibc, const double* obc,
                               const double* vi1, const double* ks, const double* psi0, const double* psi1,
                               const double* ipsi1, const double* ipsi2, const double* combo1, const double* combo2,
                               const double* m3, const double* kinv_mode_minor, AnnularStokesPlan** out);
void annular_stokes_plan_destroy(AnnularStokesPlan* p);
int annular_stokes_apply(AnnularStokesPlan* p, const cplx* in, cplx* out, cudaStream_t st);
int annular_stokes_precond(AnnularStokesPlan* p, const cplx* in, cplx* out, cudaStream_t st);
int annular_stokes_solve(AnnularStokesPlan* p, const double* fr, const double* ft, const double* irg, const double* itg,
                         const double* org, const double* otg, double tol, int restart, int maxiter, double* ur, double* ut,
                         double* pr, int* iters, double* relres, cudaStream_t st);
