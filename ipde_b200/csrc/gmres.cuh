// Right-preconditioned restarted GMRES on the device (CGS2 Arnoldi; Hessenberg/Givens on the host).
// Stands in for personal_utilities.scipy_gmres.right_gmres as called at
// ipde/annular/modified_helmholtz.py:198 and ipde/annular/stokes.py:533.
#pragma once
#include "common.cuh"

struct LinearOperator {
    virtual ~LinearOperator() {}
    virtual int apply(const cplx* in, cplx* out, cudaStream_t st) = 0;    // out = A in
    virtual int precond(const cplx* in, cplx* out, cudaStream_t st) = 0;  // out = M in
};

struct GmresWorkspace {
    int64_t len = 0;
    int restart = 0;
    cplx* V = nullptr;       // (restart+1) x len
    cplx* w = nullptr;
    cplx* z = nullptr;
    cplx* y = nullptr;       // len: preconditioned-variable solution
    cplx* hbuf = nullptr;    // device: [h (restart+1) | h2 (restart+1) | norm (1)]
    double* partial = nullptr;
    cplx* host = nullptr;    // pinned mirror of hbuf
    int nblocks = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};   // completion of the host copies of two in-flight iterations
    unsigned* ticket = nullptr;   // last-block-done counter of the fused dot-product kernel
};

int gmres_workspace_ensure(GmresWorkspace* ws, int64_t len, int restart);
void gmres_workspace_free(GmresWorkspace* ws);

// Solves A x = b.  x (out) = M y.  Returns iteration count through *iters and the final relative residual.
int gmres_solve(LinearOperator* op, GmresWorkspace* ws, const cplx* b, cplx* x, int64_t len, double tol, int restart,
                int maxiter, int* iters, double* relres, cudaStream_t st);
