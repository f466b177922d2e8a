// Direct (all-pairs) evaluation of sums of free-space Green's functions, FP64:
//     u(x_t) = sum_s G(|x_t - y_s|) q_s
// G = -log r / (2 pi)            (Laplace;   reference grid_evaluators/laplace_grid_evaluator.py:8-12)
// G = K0(k r) / (2 pi)           (modified Helmholtz; modified_helmholtz_grid_evaluator.py:8-9)
// Stokes: stokeslet velocity (1/4pi)(-log r I + r r^T / r^2) f and pressure (1/2pi) r.f / r^2
//                                (solvers/internals/stokes_save.py:29-38; pyfmmlib2d.SFMM call at internals/stokes.py:29-34)
// Stands in for pybie2d's *_Layer_Apply(backend='fly') / SFMM as called from
// ipde/solvers/internals/scalar.py:101,113, multi_boundary/scalar.py:67 and modified_helmholtz.py:63.
// q_s already includes the source quadrature weight.
//
// CUDA's math library has no double-precision K0, so it is evaluated here: power series for x <= 2,
// exp(-x)/sqrt(x) times a Chebyshev series in 4/x - 1 beyond (coefficients fitted on the host against
// scipy.special.k0e and uploaded with ipde_set_k0_table).
#include "layer.cuh"
#include <algorithm>

#include "bessel.cuh"
__constant__ double c_k0_cheb[K0_MAX_CHEB];
__constant__ double c_k1_cheb[K0_MAX_CHEB];
__constant__ int c_k0_ncheb;
__constant__ double c_k0_a[16];
__constant__ double c_k0_b[16];
__constant__ double c_k1_a[16];
__constant__ double c_k1_b[16];
// the K0/K1 tables live in __constant__ memory, i.e. once per device: the flag is per device too
static bool g_k0_ready_dev[64] = {false};
static bool k0_ready() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < 64 && g_k0_ready_dev[dev];
}
#define g_k0_ready k0_ready()
bool bessel_tables_ready() { return k0_ready(); }

int layer_set_k0_table(const double* cheb, const double* cheb1, int ncheb) {
    ARG_CHECK(ncheb >= 2 && ncheb <= K0_MAX_CHEB, "bad Chebyshev table length");
    CUDA_TRY(cudaMemcpyToSymbol(c_k0_cheb, cheb, sizeof(double) * ncheb));
    CUDA_TRY(cudaMemcpyToSymbol(c_k1_cheb, cheb1, sizeof(double) * ncheb));
    CUDA_TRY(cudaMemcpyToSymbol(c_k0_ncheb, &ncheb, sizeof(int)));
    double a[16], b[16], a1[16], b1[16];
    double fact = 1.0, H = 0.0;
    for (int m = 0; m < 16; ++m) {
        if (m > 0) { fact *= m; H += 1.0 / m; }
        a[m] = 1.0 / (fact * fact);
        b[m] = H / (fact * fact);
        const double f1 = fact * fact * (m + 1);            // m! (m+1)!
        a1[m] = 1.0 / f1;
        b1[m] = (H + H + 1.0 / (m + 1)) / (2.0 * f1);
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_k0_a, a, sizeof(a)));
    CUDA_TRY(cudaMemcpyToSymbol(c_k0_b, b, sizeof(b)));
    CUDA_TRY(cudaMemcpyToSymbol(c_k1_a, a1, sizeof(a1)));
    CUDA_TRY(cudaMemcpyToSymbol(c_k1_b, b1, sizeof(b1)));
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) g_k0_ready_dev[dev] = true;
    return IPDE_OK;
}

__global__ void k_eval_k1(const double* __restrict__ x, double* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bessel_k1(x[i], x[i] * x[i]);
}

__global__ void k_eval_k0(const double* __restrict__ x, double* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bessel_k0(x[i], x[i] * x[i]);
}

int layer_eval_k0(const double* x, double* out, int64_t n, int order, cudaStream_t st) {
    ARG_CHECK(g_k0_ready, "K0 table not set");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), 4096));
    if (order == 0) k_eval_k0<<<blocks, 256, 0, st>>>(x, out, n);
    else k_eval_k1<<<blocks, 256, 0, st>>>(x, out, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

#define LAYER_TILE 256
#define LAYER_THREADS 128

// kind 0: Laplace, 1: modified Helmholtz.  partial[chunk*ntrg + t]
template <int KIND>
__global__ void k_layer_scalar(const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ q,
                               int64_t nsrc, const double* __restrict__ tx, const double* __restrict__ ty, int64_t ntrg,
                               double kk, int64_t chunk_len, double* __restrict__ partial) {
    __shared__ double shx[LAYER_TILE], shy[LAYER_TILE], shq[LAYER_TILE];
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t s0 = blockIdx.y * chunk_len;
    const int64_t s1 = min(nsrc, s0 + chunk_len);
    const bool live = t < ntrg;
    const double x = live ? tx[t] : 0.0, y = live ? ty[t] : 0.0;
    double acc = 0.0;
    for (int64_t base = s0; base < s1; base += LAYER_TILE) {
        const int cnt = (int)min((int64_t)LAYER_TILE, s1 - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            shx[i] = sx[base + i]; shy[i] = sy[base + i]; shq[i] = q[base + i];
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const double dx = x - shx[i], dy = y - shy[i];
                const double r2 = dx * dx + dy * dy;
                if (KIND == 0) {
                    acc -= log(r2) * shq[i];
                } else {
                    const double xr = kk * sqrt(r2);
                    acc += bessel_k0(xr, kk * kk * r2) * shq[i];
                }
            }
        }
        __syncthreads();
    }
    if (live) {
        const double scale = (KIND == 0) ? 0.07957747154594767280 /* 1/(4 pi) */ : 0.15915494309189533577 /* 1/(2 pi) */;
        partial[(int64_t)blockIdx.y * ntrg + t] = acc * scale;
    }
}

// normal derivative at the targets of the same sums: d/dn_t sum_s G(t - s) q_s, (tnx, tny) the target normals
//   Laplace: -(d.n)/(2 pi r^2);  modified Helmholtz: -k K1(k r) (d.n)/(2 pi r),  d = t - s
template <int KIND>
__global__ void k_layer_scalar_dn(const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ q,
                                  int64_t nsrc, const double* __restrict__ tx, const double* __restrict__ ty,
                                  const double* __restrict__ tnx, const double* __restrict__ tny, int64_t ntrg, double kk,
                                  int64_t chunk_len, double* __restrict__ partial) {
    __shared__ double shx[LAYER_TILE], shy[LAYER_TILE], shq[LAYER_TILE];
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t s0 = blockIdx.y * chunk_len;
    const int64_t s1 = min(nsrc, s0 + chunk_len);
    const bool live = t < ntrg;
    const double x = live ? tx[t] : 0.0, y = live ? ty[t] : 0.0;
    const double nx = live ? tnx[t] : 0.0, ny = live ? tny[t] : 0.0;
    double acc = 0.0;
    for (int64_t base = s0; base < s1; base += LAYER_TILE) {
        const int cnt = (int)min((int64_t)LAYER_TILE, s1 - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            shx[i] = sx[base + i]; shy[i] = sy[base + i]; shq[i] = q[base + i];
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const double dx = x - shx[i], dy = y - shy[i];
                const double r2 = dx * dx + dy * dy;
                const double dn = dx * nx + dy * ny;
                if (KIND == 0) {
                    acc -= dn / r2 * shq[i];
                } else {
                    const double r = sqrt(r2);
                    acc -= kk * bessel_k1(kk * r, kk * kk * r2) * dn / r * shq[i];
                }
            }
        }
        __syncthreads();
    }
    if (live) partial[(int64_t)blockIdx.y * ntrg + t] = acc * 0.15915494309189533577;   // 1/(2 pi)
}

// Stokes: forces (fx, fy) already weighted; outputs u, v, p
__global__ void k_layer_stokes(const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ fx,
                               const double* __restrict__ fy, int64_t nsrc, const double* __restrict__ tx,
                               const double* __restrict__ ty, int64_t ntrg, int64_t chunk_len, double* __restrict__ partial) {
    __shared__ double shx[LAYER_TILE], shy[LAYER_TILE], shfx[LAYER_TILE], shfy[LAYER_TILE];
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t s0 = blockIdx.y * chunk_len;
    const int64_t s1 = min(nsrc, s0 + chunk_len);
    const bool live = t < ntrg;
    const double x = live ? tx[t] : 0.0, y = live ? ty[t] : 0.0;
    double au = 0.0, av = 0.0, ap = 0.0;
    for (int64_t base = s0; base < s1; base += LAYER_TILE) {
        const int cnt = (int)min((int64_t)LAYER_TILE, s1 - base);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            shx[i] = sx[base + i]; shy[i] = sy[base + i]; shfx[i] = fx[base + i]; shfy[i] = fy[base + i];
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const double dx = x - shx[i], dy = y - shy[i];
                const double r2 = dx * dx + dy * dy;
                const double ir2 = 1.0 / r2;
                const double L = -0.5 * log(r2);
                const double rf = (dx * shfx[i] + dy * shfy[i]) * ir2;
                au += L * shfx[i] + rf * dx;
                av += L * shfy[i] + rf * dy;
                ap += rf;
            }
        }
        __syncthreads();
    }
    if (live) {
        const double c4 = 0.07957747154594767280, c2 = 0.15915494309189533577;
        double* base = partial + (int64_t)blockIdx.y * 3 * ntrg;
        base[t] = au * c4;
        base[ntrg + t] = av * c4;
        base[2 * ntrg + t] = ap * c2;
    }
}

__global__ void k_layer_reduce(const double* __restrict__ partial, int nchunks, int64_t n, double* __restrict__ out,
                               int accumulate) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = accumulate ? out[i] : 0.0;
        for (int c = 0; c < nchunks; ++c) acc += partial[(int64_t)c * n + i];
        out[i] = acc;
    }
}

// per-call scratch from the stream-ordered allocator: no state shared between streams or devices
struct PartialScratch {
    double* ptr = nullptr;
    cudaStream_t st;
    explicit PartialScratch(cudaStream_t s) : st(s) {}
    int get(size_t n) { CUDA_TRY(cudaMallocAsync(&ptr, sizeof(double) * n, st)); return IPDE_OK; }
    ~PartialScratch() { if (ptr) cudaFreeAsync(ptr, st); }
};

static void choose_chunks(int64_t nsrc, int64_t ntrg, int* nchunks, int64_t* chunk_len) {
    const int64_t tblocks = cdiv(ntrg, LAYER_THREADS);
    const int64_t want = (int64_t)ipde_num_sms() * 4;
    int64_t c = 1;
    if (tblocks < want) c = std::min<int64_t>(cdiv(want, tblocks), cdiv(nsrc, LAYER_TILE));
    if (c < 1) c = 1;
    if (c > 4096) c = 4096;
    int64_t len = cdiv(cdiv(nsrc, c), LAYER_TILE) * LAYER_TILE;
    *nchunks = (int)cdiv(nsrc, len);
    *chunk_len = len;
}

int layer_apply_scalar(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                       const double* tx, const double* ty, int64_t ntrg, double* out, int accumulate, cudaStream_t st) {
    if (ntrg <= 0) return IPDE_OK;
    if (nsrc <= 0) {
        if (!accumulate) CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * ntrg, st));
        return IPDE_OK;
    }
    if (kind == 1) ARG_CHECK(g_k0_ready, "K0 table not set (call ipde_set_k0_table)");
    int nchunks; int64_t chunk_len;
    choose_chunks(nsrc, ntrg, &nchunks, &chunk_len);
    PartialScratch scratch(st);
    int rc = scratch.get((size_t)nchunks * ntrg);
    if (rc) return rc;
    double* partial = scratch.ptr;
    dim3 grid((unsigned)cdiv(ntrg, LAYER_THREADS), (unsigned)nchunks);
    if (kind == 0)
        k_layer_scalar<0><<<grid, LAYER_THREADS, 0, st>>>(sx, sy, q, nsrc, tx, ty, ntrg, k, chunk_len, partial);
    else
        k_layer_scalar<1><<<grid, LAYER_THREADS, 0, st>>>(sx, sy, q, nsrc, tx, ty, ntrg, k, chunk_len, partial);
    KERNEL_CHECK();
    k_layer_reduce<<<(int)std::max<int64_t>(1, std::min<int64_t>(cdiv(ntrg, 256), 2048)), 256, 0, st>>>(partial, nchunks, ntrg, out, accumulate);
    KERNEL_CHECK();
    return IPDE_OK;
}

int layer_apply_scalar_dn(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                          const double* tx, const double* ty, const double* tnx, const double* tny, int64_t ntrg, double* out,
                          int accumulate, cudaStream_t st) {
    if (ntrg <= 0) return IPDE_OK;
    if (nsrc <= 0) {
        if (!accumulate) CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * ntrg, st));
        return IPDE_OK;
    }
    if (kind == 1) ARG_CHECK(g_k0_ready, "K0 table not set (call ipde_set_k0_table)");
    int nchunks; int64_t chunk_len;
    choose_chunks(nsrc, ntrg, &nchunks, &chunk_len);
    PartialScratch scratch(st);
    int rc = scratch.get((size_t)nchunks * ntrg);
    if (rc) return rc;
    double* partial = scratch.ptr;
    dim3 grid((unsigned)cdiv(ntrg, LAYER_THREADS), (unsigned)nchunks);
    if (kind == 0)
        k_layer_scalar_dn<0><<<grid, LAYER_THREADS, 0, st>>>(sx, sy, q, nsrc, tx, ty, tnx, tny, ntrg, k, chunk_len, partial);
    else
        k_layer_scalar_dn<1><<<grid, LAYER_THREADS, 0, st>>>(sx, sy, q, nsrc, tx, ty, tnx, tny, ntrg, k, chunk_len, partial);
    KERNEL_CHECK();
    k_layer_reduce<<<(int)std::max<int64_t>(1, std::min<int64_t>(cdiv(ntrg, 256), 2048)), 256, 0, st>>>(partial, nchunks, ntrg, out, accumulate);
    KERNEL_CHECK();
    return IPDE_OK;
}

int layer_apply_stokes(const double* sx, const double* sy, const double* fx, const double* fy, int64_t nsrc,
                       const double* tx, const double* ty, int64_t ntrg, double* out_uvp, int accumulate, cudaStream_t st) {
    if (ntrg <= 0) return IPDE_OK;
    if (nsrc <= 0) {
        if (!accumulate) CUDA_TRY(cudaMemsetAsync(out_uvp, 0, sizeof(double) * 3 * ntrg, st));
        return IPDE_OK;
    }
    int nchunks; int64_t chunk_len;
    choose_chunks(nsrc, ntrg, &nchunks, &chunk_len);
    PartialScratch scratch(st);
    int rc = scratch.get((size_t)nchunks * 3 * ntrg);
    if (rc) return rc;
    double* partial = scratch.ptr;
    dim3 grid((unsigned)cdiv(ntrg, LAYER_THREADS), (unsigned)nchunks);
    k_layer_stokes<<<grid, LAYER_THREADS, 0, st>>>(sx, sy, fx, fy, nsrc, tx, ty, ntrg, chunk_len, partial);
    KERNEL_CHECK();
    k_layer_reduce<<<(int)std::max<int64_t>(1, std::min<int64_t>(cdiv(3 * ntrg, 256), 2048)), 256, 0, st>>>(partial, nchunks, 3 * ntrg, out_uvp, accumulate);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Dense kernel matrices for the QFS setup (ipde_b200/qfs.py): out[i*ns + j] = K(t_i, s_j) w_j with
//   layer 0 (single): G(|t - s|);   layer 1 (double): d/dn_s G = grad_s G . n_s
// for G = -log r/(2 pi) (kind 0) or K0(k r)/(2 pi) (kind 1).  Setup-time only, but O(N^2) evaluations of
// K0/K1, which is why it runs here rather than in SciPy.
template <int KIND, int LAYER>
__global__ void k_kernel_matrix(const double* __restrict__ tx, const double* __restrict__ ty, int64_t nt,
                                const double* __restrict__ sx, const double* __restrict__ sy,
                                const double* __restrict__ snx, const double* __restrict__ sny,
                                const double* __restrict__ w, int64_t ns, double kk, double* __restrict__ out) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= ns) return;
    const double xs = sx[j], ys = sy[j], wj = w[j];
    const double nxs = LAYER ? snx[j] : 0.0, nys = LAYER ? sny[j] : 0.0;
    const double c2 = 0.15915494309189533577;     // 1/(2 pi)
    for (int64_t i = blockIdx.y; i < nt; i += gridDim.y) {
        const double dx = tx[i] - xs, dy = ty[i] - ys;
        const double r2 = dx * dx + dy * dy;
        double v;
        if (LAYER == 0) {
            if (KIND == 0) v = -0.5 * log(r2) * c2;
            else v = bessel_k0(kk * sqrt(r2), kk * kk * r2) * c2;
        } else {
            const double dn = dx * nxs + dy * nys;
            if (KIND == 0) v = dn / r2 * c2;
            else {
                const double r = sqrt(r2);
                v = kk * bessel_k1(kk * r, kk * kk * r2) * dn / r * c2;
            }
        }
        out[i * ns + j] = v * wj;
    }
}

int layer_kernel_matrix(int kind, double k, int layer, const double* tx, const double* ty, int64_t nt, const double* sx,
                        const double* sy, const double* snx, const double* sny, const double* w, int64_t ns, double* out,
                        cudaStream_t st) {
    if (nt <= 0 || ns <= 0) return IPDE_OK;
    ARG_CHECK((kind == 0 || kind == 1) && (layer == 0 || layer == 1), "kind/layer must be 0 or 1");
    if (kind == 1) ARG_CHECK(g_k0_ready, "K0 table not set (call ipde_set_k0_table)");
    ARG_CHECK(layer == 0 || (snx && sny), "double layer needs source normals");
    dim3 grid((unsigned)cdiv(ns, 128), (unsigned)std::min<int64_t>(nt, 4096));
#define KM(K, L) k_kernel_matrix<K, L><<<grid, 128, 0, st>>>(tx, ty, nt, sx, sy, snx, sny, w, ns, k, out)
    if (kind == 0 && layer == 0) KM(0, 0);
    else if (kind == 0) KM(0, 1);
    else if (layer == 0) KM(1, 0);
    else KM(1, 1);
#undef KM
    KERNEL_CHECK();
    return IPDE_OK;
}
