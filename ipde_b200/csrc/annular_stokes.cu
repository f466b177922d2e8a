// Annular Stokes solver:  -mu Lap u + grad p = f, div u = 0 on a curvilinear annulus in (r, t) components,
// Dirichlet data at both radii, right-preconditioned GMRES in theta-Fourier space with the Nyquist mode dropped.
//
// Device restatement of
//   AnnularStokesSolver._apply_optim_real   ipde/annular/stokes.py:321-385   (the operator given to GMRES)
//   AnnularStokesSolver._preconditioner     ipde/annular/stokes.py:200-210, batch_matvec :51-71
//   AnnularStokesSolver.solve               ipde/annular/stokes.py:519-541
//   mfft / mifft / mifftr                   ipde/utilities.py:78-99
// Unknown vector: (3M-1, n) complex, rows [u_r (M) | u_t (M) | p (M-1)], mode-minor, the Nyquist column kept at
// zero (the reference stores ns = n-1 columns).  Equation rows: [M-2 momentum_r, ibc_r, obc_r | M-2 momentum_t,
// ibc_t, obc_t | M-1 divergence].  The per-mode (3M-1)^2 complex inverses (stokes.py:127-153) are built on the
// host (ipde_b200/annular.py) and uploaded mode-minor.
#include "annular.cuh"
#include <algorithm>

#define AST_MAX_M 40

struct AnnularStokesOp;

struct AnnularStokesPlan {
    int n = 0, M = 0;
    double mu = 1.0;
    FftPlan* fft = nullptr;
    double *D01 = nullptr, *D12 = nullptr, *R01 = nullptr, *R12 = nullptr, *R02 = nullptr, *P10 = nullptr;
    double *ibc = nullptr, *obc = nullptr, *vi1 = nullptr;   // (M), (M), first row of VI1 (M-1)
    double* ks = nullptr;                                    // n (0 at the Nyquist column)
    // geometry, real (rows, n): psi0 (M), psi1 (M-1), ipsi1 (M-1), ipsi2 (M-2), combo1, combo2, m3 (M-2 each)
    double *psi0 = nullptr, *psi1 = nullptr, *ipsi1 = nullptr, *ipsi2 = nullptr, *combo1 = nullptr, *combo2 = nullptr,
           *m3 = nullptr;
    cplx* kinv = nullptr;   // ((3M-1), (3M-1), n)
    cplx *S1 = nullptr;     // (6M-2, n): [ur, ut, p, d ur, d ut, d p] spectra -> real space
    cplx *T = nullptr;      // (2(M-1), n): (R01 d u) ipsi1 -> theta derivative
    cplx *rhs = nullptr, *sol = nullptr;
    GmresWorkspace ws;
    AnnularStokesOp* op = nullptr;
};

// S1 rows: [in (3M-1) ; i k in (3M-1)]
__global__ void k_ast_pre(const cplx* __restrict__ in, cplx* __restrict__ S1, int n, int rows, const double* __restrict__ ks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = blockIdx.y;
    const cplx v = in[(int64_t)r * n + i];
    const double k = ks[i];
    S1[(int64_t)r * n + i] = v;
    S1[(int64_t)(rows + r) * n + i] = make_double2(-k * v.y, k * v.x);
}

// T rows: (R01 d ur) ipsi1, (R01 d ut) ipsi1   (real parts of S1)
__global__ void k_ast_stageA(const cplx* __restrict__ S1, cplx* __restrict__ T, int n, int M, const double* __restrict__ R01,
                             const double* __restrict__ ipsi1) {
    extern __shared__ double shm[];
    for (int t = threadIdx.x; t < (M - 1) * M; t += blockDim.x) shm[t] = R01[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rows = 3 * M - 1;
    const cplx* dur = S1 + (int64_t)rows * n;
    const cplx* dut = dur + (int64_t)M * n;
    double a[AST_MAX_M], b[AST_MAX_M];
    for (int r = 0; r < M; ++r) { a[r] = dur[(int64_t)r * n + i].x; b[r] = dut[(int64_t)r * n + i].x; }
    for (int q = 0; q < M - 1; ++q) {
        double sa = 0.0, sb = 0.0;
        for (int r = 0; r < M; ++r) { sa = fma(shm[q * M + r], a[r], sa); sb = fma(shm[q * M + r], b[r], sb); }
        const double w = ipsi1[(int64_t)q * n + i];
        T[(int64_t)q * n + i] = make_double2(sa * w, 0.0);
        T[(int64_t)(M - 1 + q) * n + i] = make_double2(sb * w, 0.0);
    }
}

// multiply rows by i k (Nyquist column -> 0)
__global__ void k_ast_ik(cplx* __restrict__ x, int n, const double* __restrict__ ks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = (int64_t)blockIdx.y * n + i;
    const cplx v = x[o];
    const double k = ks[i];
    x[o] = make_double2(-k * v.y, k * v.x);
}

struct AstMats {
    const double *D01, *D12, *R01, *R12, *R02;
};
struct AstGeom {
    const double *psi0, *psi1, *ipsi1, *ipsi2, *combo1, *combo2, *m3;
};

// out rows (real space, complex with zero imaginary part): momentum_r (M-2) at row 0, momentum_t (M-2) at row M,
// divergence (M-1) at row 2M
__global__ void k_ast_stageB(const cplx* __restrict__ S1, const cplx* __restrict__ T, cplx* __restrict__ out, int n, int M,
                             AstMats A, AstGeom G, double mu) {
    extern __shared__ double shm[];
    double* sD01 = shm;                        // (M-1) x M
    double* sR01 = sD01 + (M - 1) * M;
    double* sD12 = sR01 + (M - 1) * M;         // (M-2) x (M-1)
    double* sR12 = sD12 + (M - 2) * (M - 1);
    double* sR02 = sR12 + (M - 2) * (M - 1);   // (M-2) x M
    for (int t = threadIdx.x; t < (M - 1) * M; t += blockDim.x) { sD01[t] = A.D01[t]; sR01[t] = A.R01[t]; }
    for (int t = threadIdx.x; t < (M - 2) * (M - 1); t += blockDim.x) { sD12[t] = A.D12[t]; sR12[t] = A.R12[t]; }
    for (int t = threadIdx.x; t < (M - 2) * M; t += blockDim.x) sR02[t] = A.R02[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rows = 3 * M - 1;
#define RE(base, r) ((base)[(int64_t)(r) * n + i].x)
    const cplx* s_ur = S1;
    const cplx* s_ut = S1 + (int64_t)M * n;
    const cplx* s_p = S1 + (int64_t)2 * M * n;
    const cplx* s_dur = S1 + (int64_t)rows * n;
    const cplx* s_dut = s_dur + (int64_t)M * n;
    const cplx* s_dp = s_dur + (int64_t)2 * M * n;
    double ur[AST_MAX_M], ut[AST_MAX_M], dur[AST_MAX_M], dut[AST_MAX_M], p[AST_MAX_M], dp[AST_MAX_M];
    for (int r = 0; r < M; ++r) { ur[r] = RE(s_ur, r); ut[r] = RE(s_ut, r); dur[r] = RE(s_dur, r); dut[r] = RE(s_dut, r); }
    for (int r = 0; r < M - 1; ++r) { p[r] = RE(s_p, r); dp[r] = RE(s_dp, r); }
    // level-1 radial derivatives on the M-1 grid, times psi1; theta-derivative terms come in through T
    double gr[AST_MAX_M], gt[AST_MAX_M], tr[AST_MAX_M], tt[AST_MAX_M], w2[AST_MAX_M], dv[AST_MAX_M];
    for (int q = 0; q < M - 1; ++q) {
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        for (int r = 0; r < M; ++r) {
            a = fma(sD01[q * M + r], ur[r], a);
            b = fma(sD01[q * M + r], ut[r], b);
            c = fma(sR01[q * M + r], dut[r], c);
            d = fma(sD01[q * M + r], ur[r] * G.psi0[(int64_t)r * n + i], d);
        }
        const double ps = G.psi1[(int64_t)q * n + i];
        gr[q] = a * ps; gt[q] = b * ps;
        w2[q] = c; dv[q] = d;
        tr[q] = RE(T, q); tt[q] = RE(T, M - 1 + q);
    }
    // divergence rows
    for (int q = 0; q < M - 1; ++q)
        out[(int64_t)(2 * M + q) * n + i] = make_double2((dv[q] + w2[q]) * G.ipsi1[(int64_t)q * n + i], 0.0);
    // momentum rows on the M-2 grid
    for (int q = 0; q < M - 2; ++q) {
        double rr_r = 0.0, rr_t = 0.0, tt_r = 0.0, tt_t = 0.0, d12p = 0.0, r12dp = 0.0;
        for (int r = 0; r < M - 1; ++r) {
            const double d = sD12[q * (M - 1) + r], rr = sR12[q * (M - 1) + r];
            rr_r = fma(d, gr[r], rr_r); rr_t = fma(d, gt[r], rr_t);
            tt_r = fma(rr, tr[r], tt_r); tt_t = fma(rr, tt[r], tt_t);
            d12p = fma(d, p[r], d12p); r12dp = fma(rr, dp[r], r12dp);
        }
        double w1r = 0.0, w1t = 0.0, rdur = 0.0, rdut = 0.0;
        for (int r = 0; r < M; ++r) {
            const double m = sR02[q * M + r];
            w1r = fma(m, ur[r], w1r); w1t = fma(m, ut[r], w1t);
            rdur = fma(m, dur[r], rdur); rdut = fma(m, dut[r], rdut);
        }
        const int64_t o = (int64_t)q * n + i;
        const double ip2 = G.ipsi2[o], c1 = G.combo1[o], c2 = G.combo2[o], m3 = G.m3[o];
        const double fr = mu * (-(rr_r + tt_r) * ip2 + rdut * c1 + w1r * c2 + w1t * m3) + d12p;
        const double ft = mu * (-(rr_t + tt_t) * ip2 - rdur * c1 + w1t * c2 - w1r * m3) + r12dp * ip2;
        out[(int64_t)q * n + i] = make_double2(fr, 0.0);
        out[(int64_t)(M + q) * n + i] = make_double2(ft, 0.0);
    }
#undef RE
}

// boundary rows (from the input spectra), mean-pressure term, Nyquist column -> 0
__global__ void k_ast_post(const cplx* __restrict__ in, cplx* __restrict__ out, int n, int M, const double* __restrict__ ibc,
                           const double* __restrict__ obc, const double* __restrict__ vi1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rows = 3 * M - 1;
    if (i == n / 2) {
        for (int r = 0; r < rows; ++r) out[(int64_t)r * n + i] = make_double2(0.0, 0.0);
        return;
    }
    for (int comp = 0; comp < 2; ++comp) {
        cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
        for (int r = 0; r < M; ++r) {
            const cplx u = in[(int64_t)(comp * M + r) * n + i];
            a.x = fma(ibc[r], u.x, a.x); a.y = fma(ibc[r], u.y, a.y);
            b.x = fma(obc[r], u.x, b.x); b.y = fma(obc[r], u.y, b.y);
        }
        out[(int64_t)(comp * M + M - 2) * n + i] = a;
        out[(int64_t)(comp * M + M - 1) * n + i] = b;
    }
    if (i == 0) {
        cplx m = make_double2(0.0, 0.0);
        for (int r = 0; r < M - 1; ++r) {
            const cplx pv = in[(int64_t)(2 * M + r) * n];
            m.x = fma(vi1[r], pv.x, m.x); m.y = fma(vi1[r], pv.y, m.y);
        }
        for (int r = 0; r < M - 1; ++r) {
            cplx v = out[(int64_t)(2 * M + r) * n];
            out[(int64_t)(2 * M + r) * n] = make_double2(v.x + m.x, v.y + m.y);
        }
    }
}

// per-mode dense complex (3M-1)^2 mat-vec; thread (mode i, output row j = blockIdx.y)
__global__ void k_ast_precond(const cplx* __restrict__ x, cplx* __restrict__ out, int n, int R, const cplx* __restrict__ kinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = blockIdx.y;
    const cplx* a = kinv + (int64_t)j * R * n + i;
    cplx acc = make_double2(0.0, 0.0);
    for (int k = 0; k < R; ++k) {
        const cplx m = a[(int64_t)k * n], v = x[(int64_t)k * n + i];
        acc.x = fma(m.x, v.x, fma(-m.y, v.y, acc.x));
        acc.y = fma(m.x, v.y, fma(m.y, v.x, acc.y));
    }
    out[(int64_t)j * n + i] = acc;
}

struct AnnularStokesOp : public LinearOperator {
    AnnularStokesPlan* p;
    explicit AnnularStokesOp(AnnularStokesPlan* plan) : p(plan) {}
    int apply(const cplx* in, cplx* out, cudaStream_t st) override {
        const int n = p->n, M = p->M, rows = 3 * M - 1;
        const int threads = 128, blocks = (n + threads - 1) / threads;
        k_ast_pre<<<dim3(blocks, rows), threads, 0, st>>>(in, p->S1, n, rows, p->ks);
        KERNEL_CHECK();
        int rc = fft_rows(p->fft, p->S1, p->S1, 2 * rows, n, n, +1, 1.0 / n, st);
        if (rc) return rc;
        k_ast_stageA<<<blocks, threads, sizeof(double) * (M - 1) * M, st>>>(p->S1, p->T, n, M, p->R01, p->ipsi1);
        KERNEL_CHECK();
        // theta derivative with the Nyquist mode dropped: mifftr(i k mfft(.))
        rc = fft_rows(p->fft, p->T, p->T, 2 * (M - 1), n, n, -1, 1.0, st);
        if (rc) return rc;
        k_ast_ik<<<dim3(blocks, 2 * (M - 1)), threads, 0, st>>>(p->T, n, p->ks);
        KERNEL_CHECK();
        rc = fft_rows(p->fft, p->T, p->T, 2 * (M - 1), n, n, +1, 1.0 / n, st);
        if (rc) return rc;
        AstMats A = {p->D01, p->D12, p->R01, p->R12, p->R02};
        AstGeom G = {p->psi0, p->psi1, p->ipsi1, p->ipsi2, p->combo1, p->combo2, p->m3};
        const size_t sh = sizeof(double) * (2 * (M - 1) * M + 2 * (M - 2) * (M - 1) + (M - 2) * M);
        k_ast_stageB<<<blocks, threads, sh, st>>>(p->S1, p->T, out, n, M, A, G, p->mu);
        KERNEL_CHECK();
        // rows M-2, M-1, 2M-2, 2M-1 of `out` are not written by stage B; transform everything (cheap) and overwrite
        rc = fft_rows(p->fft, out, out, rows, n, n, -1, 1.0, st);
        if (rc) return rc;
        k_ast_post<<<blocks, threads, 0, st>>>(in, out, n, M, p->ibc, p->obc, p->vi1);
        KERNEL_CHECK();
        return IPDE_OK;
    }
    int precond(const cplx* in, cplx* out, cudaStream_t st) override {
        const int threads = 128, blocks = (p->n + threads - 1) / threads;
        const int R = 3 * p->M - 1;
        KERNEL_TAG("k_ast_precond", (16.0 * R + 32.0) * (double)R * (double)p->n);
        k_ast_precond<<<dim3(blocks, R), threads, 0, st>>>(in, out, p->n, R, p->kinv);
        KERNEL_CHECK();
        return IPDE_OK;
    }
};

static int up(double** dst, const double* src, size_t count) {
    CUDA_TRY(cudaMalloc(dst, sizeof(double) * count));
    CUDA_TRY(cudaMemcpy(*dst, src, sizeof(double) * count, cudaMemcpyHostToDevice));
    return IPDE_OK;
}

int annular_stokes_plan_create(int n, int M, double mu, const double* D01, const double* D12, const double* R01,
                               const double* R12, const double* R02, const double* P10, const double* ibc, const double* obc,
                               const double* vi1, const double* ks, const double* psi0, const double* psi1,
                               const double* ipsi1, const double* ipsi2, const double* combo1, const double* combo2,
                               const double* m3, const double* kinv_mode_minor, AnnularStokesPlan** out) {
    ARG_CHECK(M >= 4 && M <= AST_MAX_M, "M must be in [4, 40]");
    ARG_CHECK(n % 2 == 0, "n must be even");
    AnnularStokesPlan* p = new AnnularStokesPlan();
    p->n = n; p->M = M; p->mu = mu;
    int rc = fft_plan_create(n, &p->fft);
    if (rc) { delete p; return rc; }
    const size_t R = 3 * M - 1;
#define UPL(field, src, cnt) rc = up(&p->field, src, cnt); if (rc) { annular_stokes_plan_destroy(p); return rc; }
    UPL(D01, D01, (size_t)(M - 1) * M)
    UPL(D12, D12, (size_t)(M - 2) * (M - 1))
    UPL(R01, R01, (size_t)(M - 1) * M)
    UPL(R12, R12, (size_t)(M - 2) * (M - 1))
    UPL(R02, R02, (size_t)(M - 2) * M)
    UPL(P10, P10, (size_t)M * (M - 1))
    UPL(ibc, ibc, (size_t)M)
    UPL(obc, obc, (size_t)M)
    UPL(vi1, vi1, (size_t)(M - 1))
    UPL(ks, ks, (size_t)n)
    UPL(psi0, psi0, (size_t)M * n)
    UPL(psi1, psi1, (size_t)(M - 1) * n)
    UPL(ipsi1, ipsi1, (size_t)(M - 1) * n)
    UPL(ipsi2, ipsi2, (size_t)(M - 2) * n)
    UPL(combo1, combo1, (size_t)(M - 2) * n)
    UPL(combo2, combo2, (size_t)(M - 2) * n)
    UPL(m3, m3, (size_t)(M - 2) * n)
#undef UPL
    CUDA_TRY(cudaMalloc(&p->kinv, sizeof(cplx) * R * R * n));
    CUDA_TRY(cudaMemcpy(p->kinv, kinv_mode_minor, sizeof(cplx) * R * R * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&p->S1, sizeof(cplx) * 2 * R * n));
    CUDA_TRY(cudaMalloc(&p->T, sizeof(cplx) * (size_t)2 * (M - 1) * n));
    CUDA_TRY(cudaMalloc(&p->rhs, sizeof(cplx) * R * n));
    CUDA_TRY(cudaMalloc(&p->sol, sizeof(cplx) * R * n));
    p->op = new AnnularStokesOp(p);
    *out = p;
    return IPDE_OK;
}

void annular_stokes_plan_destroy(AnnularStokesPlan* p) {
    if (!p) return;
    fft_plan_destroy(p->fft);
    cudaFree(p->D01); cudaFree(p->D12); cudaFree(p->R01); cudaFree(p->R12); cudaFree(p->R02); cudaFree(p->P10);
    cudaFree(p->ibc); cudaFree(p->obc); cudaFree(p->vi1); cudaFree(p->ks);
    cudaFree(p->psi0); cudaFree(p->psi1); cudaFree(p->ipsi1); cudaFree(p->ipsi2); cudaFree(p->combo1); cudaFree(p->combo2);
    cudaFree(p->m3); cudaFree(p->kinv); cudaFree(p->S1); cudaFree(p->T); cudaFree(p->rhs); cudaFree(p->sol);
    gmres_workspace_free(&p->ws);
    delete p->op;
    delete p;
}

int annular_stokes_apply(AnnularStokesPlan* p, const cplx* in, cplx* out, cudaStream_t st) { return p->op->apply(in, out, st); }
int annular_stokes_precond(AnnularStokesPlan* p, const cplx* in, cplx* out, cudaStream_t st) { return p->op->precond(in, out, st); }

// rhs rows: [R02 fr ; irg ; org | R02 ft ; itg ; otg | 0]
__global__ void k_ast_rhs(const double* __restrict__ fr, const double* __restrict__ ft, const double* __restrict__ irg,
                          const double* __restrict__ itg, const double* __restrict__ org, const double* __restrict__ otg,
                          cplx* __restrict__ out, int n, int M, const double* __restrict__ R02) {
    extern __shared__ double shm[];
    for (int t = threadIdx.x; t < (M - 2) * M; t += blockDim.x) shm[t] = R02[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int comp = 0; comp < 2; ++comp) {
        const double* f = comp ? ft : fr;
        double u[AST_MAX_M];
        for (int r = 0; r < M; ++r) u[r] = f[(int64_t)r * n + i];
        for (int q = 0; q < M - 2; ++q) {
            double acc = 0.0;
            for (int r = 0; r < M; ++r) acc = fma(shm[q * M + r], u[r], acc);
            out[(int64_t)(comp * M + q) * n + i] = make_double2(acc, 0.0);
        }
        const double* ig = comp ? itg : irg;
        const double* og = comp ? otg : org;
        out[(int64_t)(comp * M + M - 2) * n + i] = make_double2(ig ? ig[i] : 0.0, 0.0);
        out[(int64_t)(comp * M + M - 1) * n + i] = make_double2(og ? og[i] : 0.0, 0.0);
    }
    for (int r = 0; r < M - 1; ++r) out[(int64_t)(2 * M + r) * n + i] = make_double2(0.0, 0.0);
}

__global__ void k_ast_zero_nyquist(cplx* __restrict__ x, int n, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) x[(int64_t)r * n + n / 2] = make_double2(0.0, 0.0);
}

// ur, ut = Re ifft ; p = P10 Re ifft(ph)
__global__ void k_ast_finish(const cplx* __restrict__ sol, double* __restrict__ ur, double* __restrict__ ut,
                             double* __restrict__ pr, int n, int M, const double* __restrict__ P10) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int r = 0; r < M; ++r) {
        ur[(int64_t)r * n + i] = sol[(int64_t)r * n + i].x;
        ut[(int64_t)r * n + i] = sol[(int64_t)(M + r) * n + i].x;
    }
    double p[AST_MAX_M];
    for (int r = 0; r < M - 1; ++r) p[r] = sol[(int64_t)(2 * M + r) * n + i].x;
    for (int q = 0; q < M; ++q) {
        double acc = 0.0;
        for (int r = 0; r < M - 1; ++r) acc = fma(P10[q * (M - 1) + r], p[r], acc);
        pr[(int64_t)q * n + i] = acc;
    }
}

int annular_stokes_solve(AnnularStokesPlan* p, const double* fr, const double* ft, const double* irg, const double* itg,
                         const double* org, const double* otg, double tol, int restart, int maxiter, double* ur, double* ut,
                         double* pr, int* iters, double* relres, cudaStream_t st) {
    const int n = p->n, M = p->M, rows = 3 * M - 1;
    const int threads = 128, blocks = (n + threads - 1) / threads;
    k_ast_rhs<<<blocks, threads, sizeof(double) * (M - 2) * M, st>>>(fr, ft, irg, itg, org, otg, p->rhs, n, M, p->R02);
    KERNEL_CHECK();
    int rc = fft_rows(p->fft, p->rhs, p->rhs, rows, n, n, -1, 1.0, st);
    if (rc) return rc;
    k_ast_zero_nyquist<<<(rows + 63) / 64, 64, 0, st>>>(p->rhs, n, rows);
    KERNEL_CHECK();
    rc = gmres_solve(p->op, &p->ws, p->rhs, p->sol, (int64_t)rows * n, tol, restart, maxiter, iters, relres, st);
    if (rc) return rc;
    rc = fft_rows(p->fft, p->sol, p->sol, rows, n, n, +1, 1.0 / n, st);
    if (rc) return rc;
    k_ast_finish<<<blocks, threads, 0, st>>>(p->sol, ur, ut, pr, n, M, p->P10);
    KERNEL_CHECK();
    return IPDE_OK;
}
