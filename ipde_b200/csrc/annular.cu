// Annular (Chebyshev-Gauss in r  x  Fourier in theta) scalar solver: (k^2 - Laplacian) u = f on a curvilinear
// annulus with Dirichlet/Robin rows at both radii, solved by right-preconditioned GMRES in theta-Fourier space.
//
// Device restatement of
//   AnnularModifiedHelmholtzSolver.solve / _apply / _optim_preconditioner  ipde/annular/modified_helmholtz.py:157-203
//   fscalar_laplacian                                                       ipde/annular/modified_helmholtz.py:25-38
//   optim_batch_matvecT (numba)                                             ipde/annular/modified_helmholtz.py:68-88
//   AnnularPoissonSolver.solve                                              ipde/annular/poisson.py:16-21
// The small Chebyshev matrices (ChebyshevOperators, annular_full.py:7-50), the psi arrays (RealAnnularGeometry,
// :87-108) and the per-mode inverses (modified_helmholtz.py:143-153) are built on the host at setup and uploaded.
#include "annular.cuh"
#include <algorithm>

#define ANN_MAX_M 40

struct AnnularScalarOp;

struct AnnularPlan {
    int n = 0, M = 0;
    double k2 = 0.0;
    FftPlan* fft = nullptr;
    // device copies
    double *D01 = nullptr, *D12 = nullptr, *R01 = nullptr, *R12 = nullptr, *R02 = nullptr, *ibc = nullptr, *obc = nullptr;
    double* ks = nullptr;        // n
    double* mul1 = nullptr;      // (2(M-1), n): [psi1 ; 1/psi1]
    double* mul2 = nullptr;      // (M-2, n): 1/psi2
    double* kinv = nullptr;      // (M, M, n): kinv[(j*M + k)*n + i] = Kinv_i[j][k]
    cplx *W1 = nullptr, *W2 = nullptr, *rhs = nullptr, *sol = nullptr;
    GmresWorkspace ws;
    AnnularScalarOp* op = nullptr;
};

// The three small-matrix stages are parallel over (mode column i, output row q = blockIdx.y): one thread produces one
// output entry from the M (or M-1) entries of its column (coalesced across the threads of a block, re-read from
// L1/L2 by the other rows) -- M times more threads in flight than a thread-per-column loop.
// stage 1:  W1[0:M-1] = D01 uh ;  W1[M-1:2M-2] = R01 (i k uh)
__global__ void k_ann_stage1(const cplx* __restrict__ uh, cplx* __restrict__ W1, int n, int M, const double* __restrict__ D01,
                             const double* __restrict__ R01, const double* __restrict__ ks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = blockIdx.y;                       // 0 .. M-2
    const double* d = D01 + q * M;
    const double* r = R01 + q * M;
    cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
    for (int t = 0; t < M; ++t) {
        const cplx u = uh[(int64_t)t * n + i];
        const double dd = d[t], rr = r[t];
        a.x = fma(dd, u.x, a.x); a.y = fma(dd, u.y, a.y);
        b.x = fma(rr, u.x, b.x); b.y = fma(rr, u.y, b.y);
    }
    const double k = ks[i];
    W1[(int64_t)q * n + i] = a;
    W1[(int64_t)(M - 1 + q) * n + i] = make_double2(-k * b.y, k * b.x);
}

// stage 2:  W2 = D12 A + R12 (i k B),  A = W1[0:M-1], B = W1[M-1:2M-2] (after the variable-coefficient multiplies)
__global__ void k_ann_stage2(const cplx* __restrict__ W1, cplx* __restrict__ W2, int n, int M, const double* __restrict__ D12,
                             const double* __restrict__ R12, const double* __restrict__ ks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = blockIdx.y;                       // 0 .. M-3
    const double* d = D12 + q * (M - 1);
    const double* r = R12 + q * (M - 1);
    const double k = ks[i];
    cplx acc = make_double2(0.0, 0.0);
    for (int t = 0; t < M - 1; ++t) {
        const cplx A = W1[(int64_t)t * n + i];
        const cplx b = W1[(int64_t)(M - 1 + t) * n + i];
        const double dd = d[t], rr = r[t];
        acc.x = fma(dd, A.x, fma(rr, -k * b.y, acc.x));
        acc.y = fma(dd, A.y, fma(rr, k * b.x, acc.y));
    }
    W2[(int64_t)q * n + i] = acc;
}

// stage 3:  out[0:M-2] = k^2 R02 uh - W2 ; out[M-2] = ibc . uh ; out[M-1] = obc . uh
__global__ void k_ann_stage3(const cplx* __restrict__ uh, const cplx* __restrict__ W2, cplx* __restrict__ out, int n, int M,
                             const double* __restrict__ R02, const double* __restrict__ ibc, const double* __restrict__ obc,
                             double k2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int q = blockIdx.y;                       // 0 .. M-1
    const double* row = (q < M - 2) ? R02 + q * M : (q == M - 2 ? ibc : obc);
    cplx acc = make_double2(0.0, 0.0);
    for (int t = 0; t < M; ++t) {
        const cplx u = uh[(int64_t)t * n + i];
        const double rr = row[t];
        acc.x = fma(rr, u.x, acc.x); acc.y = fma(rr, u.y, acc.y);
    }
    if (q < M - 2) {
        const cplx l = W2[(int64_t)q * n + i];
        acc = make_double2(k2 * acc.x - l.x, k2 * acc.y - l.y);
    }
    out[(int64_t)q * n + i] = acc;
}

// per-mode dense M x M mat-vec on mode-minor vectors; thread (mode i, output row j = blockIdx.y)
__global__ void k_ann_precond(const cplx* __restrict__ x, cplx* __restrict__ out, int n, int M, const double* __restrict__ kinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = blockIdx.y;
    const double* a = kinv + (int64_t)j * M * n + i;
    double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
    int k = 0;
    for (; k + 1 < M; k += 2) {
        const double a0 = a[(int64_t)k * n], a1 = a[(int64_t)(k + 1) * n];
        const cplx v0 = x[(int64_t)k * n + i], v1 = x[(int64_t)(k + 1) * n + i];
        ax = fma(a0, v0.x, ax); ay = fma(a0, v0.y, ay);
        bx = fma(a1, v1.x, bx); by = fma(a1, v1.y, by);
    }
    if (k < M) {
        const double a0 = a[(int64_t)k * n];
        const cplx v0 = x[(int64_t)k * n + i];
        ax = fma(a0, v0.x, ax); ay = fma(a0, v0.y, ay);
    }
    out[(int64_t)j * n + i] = make_double2(ax + bx, ay + by);
}

// rhs rows: [sign * R02 f ; ig ; og] as complex (ready for the theta-FFT)
__global__ void k_ann_rhs(const double* __restrict__ f, const double* __restrict__ ig, const double* __restrict__ og,
                          cplx* __restrict__ out, int n, int M, const double* __restrict__ R02, double sign) {
    extern __shared__ double shm[];
    for (int t = threadIdx.x; t < (M - 2) * M; t += blockDim.x) shm[t] = R02[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u[ANN_MAX_M];
    for (int r = 0; r < M; ++r) u[r] = f[(int64_t)r * n + i];
    for (int q = 0; q < M - 2; ++q) {
        double acc = 0.0;
        for (int r = 0; r < M; ++r) acc += shm[q * M + r] * u[r];
        out[(int64_t)q * n + i] = make_double2(sign * acc, 0.0);
    }
    out[(int64_t)(M - 2) * n + i] = make_double2(ig ? ig[i] : 0.0, 0.0);
    out[(int64_t)(M - 1) * n + i] = make_double2(og ? og[i] : 0.0, 0.0);
}

__global__ void k_take_real(const cplx* __restrict__ in, double* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i].x;
}

struct AnnularScalarOp : public LinearOperator {
    AnnularPlan* p;
    explicit AnnularScalarOp(AnnularPlan* plan) : p(plan) {}
    int apply(const cplx* in, cplx* out, cudaStream_t st) override {
        const int n = p->n, M = p->M;
        const int threads = 128, blocks = (n + threads - 1) / threads;
        k_ann_stage1<<<dim3(blocks, M - 1), threads, 0, st>>>(in, p->W1, n, M, p->D01, p->R01, p->ks);
        KERNEL_CHECK();
        int rc = fft_mul_fft_rows(p->fft, p->W1, p->W1, p->mul1, 2 * (M - 1), st);
        if (rc) return rc;
        k_ann_stage2<<<dim3(blocks, M - 2), threads, 0, st>>>(p->W1, p->W2, n, M, p->D12, p->R12, p->ks);
        KERNEL_CHECK();
        rc = fft_mul_fft_rows(p->fft, p->W2, p->W2, p->mul2, M - 2, st);
        if (rc) return rc;
        k_ann_stage3<<<dim3(blocks, M), threads, 0, st>>>(in, p->W2, out, n, M, p->R02, p->ibc, p->obc, p->k2);
        KERNEL_CHECK();
        return IPDE_OK;
    }
    int precond(const cplx* in, cplx* out, cudaStream_t st) override {
        const int threads = 128, blocks = (p->n + threads - 1) / threads;
        KERNEL_TAG("k_ann_precond", (8.0 * p->M + 32.0) * (double)p->M * (double)p->n);
        k_ann_precond<<<dim3(blocks, p->M), threads, 0, st>>>(in, out, p->n, p->M, p->kinv);
        KERNEL_CHECK();
        return IPDE_OK;
    }
};

static int upload(double** dst, const double* src, size_t count) {
    CUDA_TRY(cudaMalloc(dst, sizeof(double) * count));
    CUDA_TRY(cudaMemcpy(*dst, src, sizeof(double) * count, cudaMemcpyHostToDevice));
    return IPDE_OK;
}

int annular_plan_create(int n, int M, double k, const double* D01, const double* D12, const double* R01, const double* R12,
                        const double* R02, const double* ibc, const double* obc, const double* ks, const double* psi1,
                        const double* ipsi1, const double* ipsi2, const double* kinv_mode_minor, AnnularPlan** out) {
    ARG_CHECK(M >= 4 && M <= ANN_MAX_M, "M must be in [4, 40]");
    AnnularPlan* p = new AnnularPlan();
    p->n = n; p->M = M; p->k2 = k * k;
    int rc = fft_plan_create(n, &p->fft);
    if (rc) { delete p; return rc; }
#define UP(field, src, cnt) rc = upload(&p->field, src, cnt); if (rc) { annular_plan_destroy(p); return rc; }
    UP(D01, D01, (size_t)(M - 1) * M)
    UP(D12, D12, (size_t)(M - 2) * (M - 1))
    UP(R01, R01, (size_t)(M - 1) * M)
    UP(R12, R12, (size_t)(M - 2) * (M - 1))
    UP(R02, R02, (size_t)(M - 2) * M)
    UP(ibc, ibc, (size_t)M)
    UP(obc, obc, (size_t)M)
    UP(ks, ks, (size_t)n)
    UP(mul2, ipsi2, (size_t)(M - 2) * n)
    UP(kinv, kinv_mode_minor, (size_t)M * M * n)
#undef UP
    CUDA_TRY(cudaMalloc(&p->mul1, sizeof(double) * (size_t)2 * (M - 1) * n));
    CUDA_TRY(cudaMemcpy(p->mul1, psi1, sizeof(double) * (size_t)(M - 1) * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->mul1 + (size_t)(M - 1) * n, ipsi1, sizeof(double) * (size_t)(M - 1) * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&p->W1, sizeof(cplx) * (size_t)2 * (M - 1) * n));
    CUDA_TRY(cudaMalloc(&p->W2, sizeof(cplx) * (size_t)(M - 2) * n));
    CUDA_TRY(cudaMalloc(&p->rhs, sizeof(cplx) * (size_t)M * n));
    CUDA_TRY(cudaMalloc(&p->sol, sizeof(cplx) * (size_t)M * n));
    p->op = new AnnularScalarOp(p);
    *out = p;
    return IPDE_OK;
}

void annular_plan_destroy(AnnularPlan* p) {
    if (!p) return;
    fft_plan_destroy(p->fft);
    cudaFree(p->D01); cudaFree(p->D12); cudaFree(p->R01); cudaFree(p->R12); cudaFree(p->R02);
    cudaFree(p->ibc); cudaFree(p->obc); cudaFree(p->ks); cudaFree(p->mul1); cudaFree(p->mul2); cudaFree(p->kinv);
    cudaFree(p->W1); cudaFree(p->W2); cudaFree(p->rhs); cudaFree(p->sol);
    gmres_workspace_free(&p->ws);
    delete p->op;
    delete p;
}

int annular_apply(AnnularPlan* p, const cplx* in, cplx* out, cudaStream_t st) { return p->op->apply(in, out, st); }
int annular_precond(AnnularPlan* p, const cplx* in, cplx* out, cudaStream_t st) { return p->op->precond(in, out, st); }

int annular_solve(AnnularPlan* p, const double* f, const double* ig, const double* og, double f_sign, double tol,
                  int restart, int maxiter, double* u_out, int* iters, double* relres, cudaStream_t st) {
    const int n = p->n, M = p->M;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    k_ann_rhs<<<blocks, threads, sizeof(double) * (M - 2) * M, st>>>(f, ig, og, p->rhs, n, M, p->R02, f_sign);
    KERNEL_CHECK();
    int rc = fft_rows(p->fft, p->rhs, p->rhs, M, n, n, -1, 1.0, st);
    if (rc) return rc;
    rc = gmres_solve(p->op, &p->ws, p->rhs, p->sol, (int64_t)M * n, tol, restart, maxiter, iters, relres, st);
    if (rc) return rc;
    rc = fft_rows(p->fft, p->sol, p->sol, M, n, n, +1, 1.0 / n, st);
    if (rc) return rc;
    const int64_t tot = (int64_t)M * n;
    k_take_real<<<(int)std::min<int64_t>(cdiv(tot, 256), 1024), 256, 0, st>>>(p->sol, u_out, tot);
    KERNEL_CHECK();
    return IPDE_OK;
}
