// Spectral-Ewald free-space evaluator:  u(x_t) = sum_s G(|x_t - y_s|) q_s  for G = -log r/(2 pi) or
// K0(kappa r)/(2 pi), from O(1e4..1e5) sources on curves to (a) nodes of the uniform grid and (b) off-grid
// points, in FP64 to ~1e-14, in O(S sw^2 + P^2 log P) instead of O(S T).
//
// Replaces the reference's Layer_Apply / Grid_Evaluator calls on the per-solve path:
//   ScalarSolver.evaluate_to_grid_pnai           ipde/solvers/multi_boundary/scalar.py:63-71
//   ScalarHelper.correct (sources -> radial pts) ipde/solvers/internals/scalar.py:95-116
//   ScalarGridBackend.ewald_local_freespace, ScalarFreespaceGridEvaluator.__call__
//                                                ipde/grid_evaluators/scalar_grid_evaluator.py:188-231,299-307
// Mathematics and host-side preprocessing: ipde_b200/ewald.py.  Stages of one application:
//   1. k_ewald_tiles     one CTA per 16x16 tile of nodes that has a source within rc: every thread owns a node and
//                        loops over the tile's source list (staged through shared memory): Gaussian spread S_w and
//                        the local sum of G_loc; no atomics, results are deterministic.
//   2. 2-D real FFT of S_w (our Stockham FFT, grid.cu), multiplication by the truncated kernel, inverse FFT(s).
//   3. k_ewald_grid_out  out[pos] += U[node] + local[node] for the on-grid targets.
//   4. k_ewald_points    one warp per off-grid target: Gaussian-window gather of H (31 x 31 nodes, rows coalesced)
//                        plus the local sum over the sources of the target's tile.
#include "ewald.cuh"
#include <algorithm>

static const double TWO_PI = 6.283185307179586476925286766559;
#define EW_TILE 16
#define EW_CHUNK 128
#define EW_MAX_TAB (64 * 16)

struct EwaldPlan {
    int kind = 0;
    double kappa = 0, xi = 0, h = 0, rc2 = 0, rw2 = 0, x0 = 0, y0 = 0;
    int px = 0, py = 0, m = 0;
    GridPlan* fft = nullptr;
    int64_t nsrc = 0;
    double *sx = nullptr, *sy = nullptr;
    int nint = 0, ndeg = 0, nlog = 0;
    double *tabS = nullptr, *tabA = nullptr;
    double *tabdS = nullptr, *tabdA = nullptr;      // d/dz tables (gradient evaluation)
    double *Lcx = nullptr, *Lcy = nullptr;          // local gradient sums per active-tile node
    double *pv = nullptr;                           // (3, n_pt): value, d/dx, d/dy at the point targets
    int ntiles = 0, ntx = 0, nty = 0;
    int *tile_ix = nullptr, *tile_iy = nullptr, *tile_ptr = nullptr, *tile_src = nullptr, *tile_map = nullptr;
    double *multH = nullptr, *multU = nullptr;
    double *Sw = nullptr, *R = nullptr, *Lc = nullptr;
    cplx *spec = nullptr, *specw = nullptr;
    int64_t n_og = 0;
    int64_t *og_lin = nullptr, *og_out = nullptr, *og_loc = nullptr;
    unsigned* tile_mask = nullptr;
    int in_lo = 0, in_hi = 0;            // rows of S_w that can be nonzero (active tiles)
    int in_chi = 0;                      // columns of S_w beyond this one are zero (active tiles)
    int grid_chi = 0;                    // grid targets only in columns below this one
    int grid_lo = 0, grid_hi = 0;        // rows holding grid targets
    int pts_lo = 0, pts_hi = 0;          // rows touched by the gather windows of the point targets   // 256 bits per tile: which nodes are grid targets (need the local sum)
    int64_t n_pt = 0;
    double *ptx = nullptr, *pty = nullptr;
};

struct EwaldTab {
    const double* S;
    const double* A;
    int nint, ndeg, nlog;
    double inv_dz;
};

// G_loc(r), z = r^2 < rc^2, from the piecewise Chebyshev tables (shared memory copies)
__device__ __forceinline__ double ewald_gloc(double z, const double* __restrict__ sS, const double* __restrict__ sA,
                                             int nint, int ndeg, int nlog, double inv_dz) {
    const double u = z * inv_dz;
    int i = (int)u;
    if (i > nint - 1) i = nint - 1;
    const double t = 2.0 * (u - (double)i) - 1.0, t2 = 2.0 * t;
    const double* c = sS + i * ndeg;
    double b0 = 0.0, b1 = 0.0;
    for (int k = ndeg - 1; k >= 1; --k) {
        const double tmp = fma(t2, b0, c[k]) - b1;
        b1 = b0; b0 = tmp;
    }
    double s = fma(t, b0, c[0]) - b1;
    if (i < nlog) {
        const double* a = sA + i * ndeg;
        b0 = 0.0; b1 = 0.0;
        for (int k = ndeg - 1; k >= 1; --k) {
            const double tmp = fma(t2, b0, a[k]) - b1;
            b1 = b0; b0 = tmp;
        }
        s -= (fma(t, b0, a[0]) - b1) * log(z);
    }
    return s;
}

// The Gaussian window is separable: exp(-2 xi^2 |x - y_s|^2) = e_x(s, row) * e_y(s, col).  One pass fills the 2 * 16
// factors of every source of the chunk (8 exp per thread instead of one per node-source pair); a factor is zero beyond
// the window half width along its axis (square window: the corners it adds to the radial one are < 6e-15 of the peak).
#define EW_SCHUNK 64
__device__ __forceinline__ void ewald_window_factors(const double* shx, const double* shy, int cnt, double tx0, double ty0,
                                                     double h, double rw2, double two_xi2, double (*shex)[EW_TILE],
                                                     double (*shey)[EW_TILE], int tid) {
    for (int idx = tid; idx < cnt * 2 * EW_TILE; idx += EW_TILE * EW_TILE) {
        const int s = idx / (2 * EW_TILE), r = idx % (2 * EW_TILE);
        if (r < EW_TILE) {
            const double d = tx0 + r * h - shx[s];
            shex[s][r] = (d * d < rw2) ? exp(-two_xi2 * d * d) : 0.0;
        } else {
            const double d = ty0 + (r - EW_TILE) * h - shy[s];
            shey[s][r - EW_TILE] = (d * d < rw2) ? exp(-two_xi2 * d * d) : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_TILE* EW_TILE)
k_ewald_tiles(const int* __restrict__ tile_ix, const int* __restrict__ tile_iy, const int* __restrict__ tile_ptr,
              const int* __restrict__ tile_src, const double* __restrict__ sx, const double* __restrict__ sy,
              const double* __restrict__ q, EwaldTab T, double x0, double y0, double h, double rc2, double rw2,
              double two_xi2, double norm, int px, int py, double* __restrict__ Sw, double* __restrict__ Lc,
              const unsigned* __restrict__ tile_mask, int want_local) {
    __shared__ double shx[EW_SCHUNK], shy[EW_SCHUNK], shq[EW_SCHUNK];
    __shared__ double shex[EW_SCHUNK][EW_TILE], shey[EW_SCHUNK][EW_TILE];
    __shared__ double sS[EW_MAX_TAB], sA[EW_MAX_TAB];
    const int tid = threadIdx.y * EW_TILE + threadIdx.x;
    const int ntab = T.nint * T.ndeg;
    for (int i = tid; i < ntab; i += EW_TILE * EW_TILE) { sS[i] = T.S[i]; sA[i] = (i < T.nlog * T.ndeg) ? T.A[i] : 0.0; }
    const int t = blockIdx.x;
    const int ix = tile_ix[t] * EW_TILE + threadIdx.y;
    const int iy = tile_iy[t] * EW_TILE + threadIdx.x;
    const double x = x0 + ix * h, y = y0 + iy * h;
    const int s0 = tile_ptr[t], s1 = tile_ptr[t + 1];
    double spread = 0.0, loc = 0.0;
    // the local sum is only needed at nodes that are grid targets
    if (want_local && tile_mask) want_local = (tile_mask[t * 8 + (tid >> 5)] >> (tid & 31)) & 1u;
    for (int base = s0; base < s1; base += EW_SCHUNK) {
        const int cnt = min(EW_SCHUNK, s1 - base);
        __syncthreads();
        if (tid < cnt) {
            const int s = tile_src[base + tid];
            shx[tid] = sx[s]; shy[tid] = sy[s]; shq[tid] = q[s];
        }
        __syncthreads();
        ewald_window_factors(shx, shy, cnt, x0 + tile_ix[t] * EW_TILE * h, y0 + tile_iy[t] * EW_TILE * h, h, rw2, two_xi2,
                             shex, shey, tid);
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            spread = fma(shq[i] * shex[i][threadIdx.y], shey[i][threadIdx.x], spread);
            if (want_local) {
                const double dx = x - shx[i], dy = y - shy[i];
                const double d2 = dx * dx + dy * dy;
                if (d2 < rc2 && d2 > 0.0)
                    loc = fma(shq[i], ewald_gloc(d2, sS, sA, T.nint, T.ndeg, T.nlog, T.inv_dz), loc);
            }
        }
    }
    if (ix < px && iy < py) Sw[(int64_t)ix * py + iy] = spread * norm;
    if (want_local) Lc[(int64_t)t * (EW_TILE * EW_TILE) + tid] = loc;
}

__global__ void k_ewald_build_mask(const int64_t* __restrict__ loc, int64_t n, unsigned* __restrict__ mask) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = loc[i];
        if (l >= 0) atomicOr(&mask[l >> 5], 1u << (unsigned)(l & 31));
    }
}

// multiplier tables on the half spectrum (plan row order)
__device__ __forceinline__ int ew_freq_of_row(int pos, int n1, int n2) {
    if (n2 == 1) return pos;
    return pos / n2 + n1 * (pos % n2);
}
__global__ void k_ewald_mult(double* __restrict__ multH, double* __restrict__ multU, int px, int py, int64_t pitch, int n1,
                             int n2, double h, int kind, double kappa, double L, double c0, double c1, double inv8xi2) {
    const int64_t total = (int64_t)px * pitch;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int pos = (int)(t / pitch), j = (int)(t % pitch);
        const int i = ew_freq_of_row(pos, n1, n2);
        const int si = (i < (px + 1) / 2) ? i : i - px;
        const double kx = TWO_PI * (double)si / ((double)px * h), ky = TWO_PI * (double)j / ((double)py * h);
        const double k2 = kx * kx + ky * ky, k = sqrt(k2);
        double g;
        if (kind == 0) {
            if (si == 0 && j == 0) g = c1;
            else g = (1.0 - j0(L * k)) / k2 - c0 * j1(L * k) / k;
        } else {
            g = (1.0 + c0 * k * j1(L * k) - c1 * j0(L * k)) / (k2 + kappa * kappa);
        }
        multH[t] = g;
        multU[t] = g * exp(-k2 * inv8xi2);
    }
}

__global__ void k_ewald_apply_mult(const cplx* __restrict__ spec, const double* __restrict__ mult, cplx* __restrict__ out,
                                   int64_t total) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        out[t] = cscale(spec[t], mult[t]);
}

__global__ void k_ewald_grid_out(const double* __restrict__ U, const double* __restrict__ Lc, const int64_t* __restrict__ lin,
                                 const int64_t* __restrict__ loc, const int64_t* __restrict__ pos, int64_t n,
                                 double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = U[lin[i]];
        const int64_t l = loc[i];
        if (l >= 0) v += Lc[l];
        out[pos[i]] += v;
    }
}

// one warp per off-grid target
__global__ void k_ewald_points(const double* __restrict__ H, const double* __restrict__ tx, const double* __restrict__ ty,
                               int64_t npts, const int* __restrict__ tile_map, int ntx, int nty,
                               const int* __restrict__ tile_ptr, const int* __restrict__ tile_src,
                               const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ q,
                               EwaldTab T, double x0, double y0, double h, double rc2, double two_xi2, double wnorm, int m,
                               int px, int py, double* __restrict__ out, int accumulate) {
    __shared__ double sS[EW_MAX_TAB], sA[EW_MAX_TAB];
    const int ntab = T.nint * T.ndeg;
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) { sS[i] = T.S[i]; sA[i] = (i < T.nlog * T.ndeg) ? T.A[i] : 0.0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int W = 2 * m + 1;      // <= 32 (checked at plan creation)
    for (int64_t p = warp; p < npts; p += nwarps) {
        const double x = tx[p], y = ty[p];
        const double X = (x - x0) / h, Y = (y - y0) / h;
        const int i0 = (int)ceil(X - m), j0 = (int)ceil(Y - m);
        // lane b: column weight wy_b and (same index) row weight wx_b
        double wy = 0.0, wx = 0.0;
        if (lane < W) {
            const double ddx = ((double)(i0 + lane) - X) * h, ddy = ((double)(j0 + lane) - Y) * h;
            wx = exp(-two_xi2 * ddx * ddx);
            wy = exp(-two_xi2 * ddy * ddy);
        }
        double acc = 0.0;
        const bool in_cols = (lane < W) && (j0 + lane >= 0) && (j0 + lane < py);
        for (int a = 0; a < W; ++a) {
            const double wa = __shfl_sync(0xffffffffu, wx, a);
            const int ii = i0 + a;
            if (in_cols && ii >= 0 && ii < px) acc = fma(wa, H[(int64_t)ii * py + j0 + lane], acc);
        }
        acc *= wy * wnorm;
        // local part: sources of the tile containing the target
        const int fxi = (int)floor(X), fyi = (int)floor(Y);
        const int tix = fxi / EW_TILE, tiy = fyi / EW_TILE;
        if (fxi >= 0 && tix < ntx && fyi >= 0 && tiy < nty) {
            const int t = tile_map[tix * nty + tiy];
            if (t >= 0) {
                const int s0 = tile_ptr[t], s1 = tile_ptr[t + 1];
                for (int s = s0 + lane; s < s1; s += 32) {
                    const int k = tile_src[s];
                    const double dx = x - sx[k], dy = y - sy[k];
                    const double d2 = dx * dx + dy * dy;
                    if (d2 < rc2 && d2 > 0.0) acc = fma(q[k], ewald_gloc(d2, sS, sA, T.nint, T.ndeg, T.nlog, T.inv_dz), acc);
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[p] = accumulate ? out[p] + acc : acc;
    }
}

// ------------------------------------------------------------------------------------------
template <typename Tp>
static int upload_arr(Tp** dst, const Tp* src, size_t count, cudaMemcpyKind kind) {
    *dst = nullptr;
    if (count == 0) return IPDE_OK;
    CUDA_TRY(cudaMalloc(dst, sizeof(Tp) * count));
    CUDA_TRY(cudaMemcpy(*dst, src, sizeof(Tp) * count, kind));
    return IPDE_OK;
}

int ewald_plan_create(int kind, double kappa, double xi, double h, double rc, int m, double x0, double y0, int px, int py,
                      double L, double c0, double c1, int64_t nsrc, const double* sx_dev, const double* sy_dev, int nint,
                      int ndeg, int nlog, const double* tabS_host, const double* tabA_host, int ntiles,
                      const int* tile_ix_host, const int* tile_iy_host, const int* tile_ptr_host,
                      const int* tile_src_host, const int* tile_map_host, EwaldPlan** out) {
    ARG_CHECK(kind == 0 || kind == 1, "kind must be 0 (Laplace) or 1 (modified Helmholtz)");
    ARG_CHECK(2 * m + 1 <= 32, "window wider than a warp");
    ARG_CHECK(nint * ndeg <= EW_MAX_TAB && nlog <= nint, "local-kernel table too large");
    ARG_CHECK(px % 2 == 0 && py % 2 == 0, "FFT box sizes must be even");
    EwaldPlan* p = new EwaldPlan();
    p->kind = kind; p->kappa = kappa; p->xi = xi; p->h = h; p->rc2 = rc * rc; p->m = m;
    p->rw2 = (m * h) * (m * h);
    p->x0 = x0; p->y0 = y0; p->px = px; p->py = py;
    p->nsrc = nsrc; p->nint = nint; p->ndeg = ndeg; p->nlog = nlog;
    p->ntiles = ntiles;
    p->ntx = (px + EW_TILE - 1) / EW_TILE; p->nty = (py + EW_TILE - 1) / EW_TILE;
    int rc_ = grid_plan_create(px, py, h, h, &p->fft);
    if (rc_) { delete p; return rc_; }
#define EW_UP(call) rc_ = (call); if (rc_) { ewald_plan_destroy(p); return rc_; }
    EW_UP(upload_arr(&p->sx, sx_dev, (size_t)nsrc, cudaMemcpyDeviceToDevice))
    EW_UP(upload_arr(&p->sy, sy_dev, (size_t)nsrc, cudaMemcpyDeviceToDevice))
    EW_UP(upload_arr(&p->tabS, tabS_host, (size_t)nint * ndeg, cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tabA, tabA_host, (size_t)nint * ndeg, cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tile_ix, tile_ix_host, (size_t)ntiles, cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tile_iy, tile_iy_host, (size_t)ntiles, cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tile_ptr, tile_ptr_host, (size_t)ntiles + 1, cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tile_src, tile_src_host, (size_t)tile_ptr_host[ntiles], cudaMemcpyHostToDevice))
    EW_UP(upload_arr(&p->tile_map, tile_map_host, (size_t)p->ntx * p->nty, cudaMemcpyHostToDevice))
#undef EW_UP
    p->in_lo = px; p->in_hi = 0; p->in_chi = 0;
    for (int t = 0; t < ntiles; ++t) {
        p->in_lo = std::min(p->in_lo, tile_ix_host[t] * EW_TILE);
        p->in_hi = std::max(p->in_hi, std::min(px, (tile_ix_host[t] + 1) * EW_TILE));
        p->in_chi = std::max(p->in_chi, std::min(py, (tile_iy_host[t] + 1) * EW_TILE));
    }
    if (p->in_hi <= p->in_lo) { p->in_lo = 0; p->in_hi = px; }
    if (p->in_chi <= 0) p->in_chi = py;
    p->grid_chi = py;
    p->grid_lo = 0; p->grid_hi = px; p->pts_lo = 0; p->pts_hi = px;
    const size_t nreal = (size_t)px * py, nspec = (size_t)px * (py / 2 + 1);
    CUDA_TRY(cudaMalloc(&p->Sw, sizeof(double) * nreal));
    CUDA_TRY(cudaMemset(p->Sw, 0, sizeof(double) * nreal));
    CUDA_TRY(cudaMalloc(&p->R, sizeof(double) * nreal));
    CUDA_TRY(cudaMalloc(&p->Lc, sizeof(double) * (size_t)std::max(1, ntiles) * EW_TILE * EW_TILE));
    CUDA_TRY(cudaMalloc(&p->spec, sizeof(cplx) * nspec));
    CUDA_TRY(cudaMalloc(&p->specw, sizeof(cplx) * nspec));
    CUDA_TRY(cudaMalloc(&p->multH, sizeof(double) * nspec));
    CUDA_TRY(cudaMalloc(&p->multU, sizeof(double) * nspec));
    const int blocks = (int)std::min<int64_t>(cdiv((int64_t)nspec, 256), (int64_t)ipde_num_sms() * 8);
    k_ewald_mult<<<blocks, 256>>>(p->multH, p->multU, px, py, py / 2 + 1, p->fft->n1, p->fft->n2, h, kind, kappa, L, c0, c1,
                                  1.0 / (8.0 * xi * xi));
    KERNEL_CHECK();
    CUDA_TRY(cudaDeviceSynchronize());
    *out = p;
    return IPDE_OK;
}

void ewald_plan_destroy(EwaldPlan* p) {
    if (!p) return;
    grid_plan_destroy(p->fft);
    cudaFree(p->sx); cudaFree(p->sy); cudaFree(p->tabS); cudaFree(p->tabA);
    cudaFree(p->tile_ix); cudaFree(p->tile_iy); cudaFree(p->tile_ptr); cudaFree(p->tile_src); cudaFree(p->tile_map);
    cudaFree(p->multH); cudaFree(p->multU); cudaFree(p->Sw); cudaFree(p->R); cudaFree(p->Lc);
    cudaFree(p->spec); cudaFree(p->specw);
    cudaFree(p->og_lin); cudaFree(p->og_out); cudaFree(p->og_loc); cudaFree(p->ptx); cudaFree(p->pty);
    cudaFree(p->tile_mask);
    cudaFree(p->tabdS); cudaFree(p->tabdA); cudaFree(p->Lcx); cudaFree(p->Lcy); cudaFree(p->pv);
    delete p;
}

int ewald_set_grid_targets(EwaldPlan* p, int64_t n, const int64_t* lin_dev, const int64_t* loc_dev, const int64_t* out_pos_dev) {
    cudaFree(p->og_lin); cudaFree(p->og_out); cudaFree(p->og_loc);
    p->og_lin = p->og_out = p->og_loc = nullptr;
    p->n_og = n;
    int rc = upload_arr(&p->og_lin, lin_dev, (size_t)n, cudaMemcpyDeviceToDevice);
    if (!rc) rc = upload_arr(&p->og_loc, loc_dev, (size_t)n, cudaMemcpyDeviceToDevice);
    if (!rc) rc = upload_arr(&p->og_out, out_pos_dev, (size_t)n, cudaMemcpyDeviceToDevice);
    if (rc) return rc;
    p->grid_lo = 0; p->grid_hi = p->px;
    if (n > 0) {
        std::vector<int64_t> h((size_t)n);
        CUDA_TRY(cudaMemcpy(h.data(), lin_dev, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost));
        int64_t lo = h[0], hi = h[0], chi = 0;
        for (int64_t i = 0; i < n; ++i) {
            lo = std::min(lo, h[i]); hi = std::max(hi, h[i]);
            chi = std::max(chi, h[i] % p->py);
        }
        p->grid_lo = (int)(lo / p->py);
        p->grid_hi = (int)(hi / p->py) + 1;
        p->grid_chi = (int)chi + 1;
    }
    cudaFree(p->tile_mask);
    p->tile_mask = nullptr;
    if (n > 0 && p->ntiles > 0) {
        CUDA_TRY(cudaMalloc(&p->tile_mask, sizeof(unsigned) * 8 * (size_t)p->ntiles));
        CUDA_TRY(cudaMemset(p->tile_mask, 0, sizeof(unsigned) * 8 * (size_t)p->ntiles));
        k_ewald_build_mask<<<(int)std::min<int64_t>(cdiv(n, 256), 4096), 256>>>(p->og_loc, n, p->tile_mask);
        KERNEL_CHECK();
        CUDA_TRY(cudaDeviceSynchronize());
    }
    return IPDE_OK;
}

int ewald_set_point_targets(EwaldPlan* p, int64_t n, const double* tx_dev, const double* ty_dev) {
    cudaFree(p->ptx); cudaFree(p->pty);
    p->ptx = p->pty = nullptr;
    p->n_pt = n;
    int rc = upload_arr(&p->ptx, tx_dev, (size_t)n, cudaMemcpyDeviceToDevice);
    if (!rc) rc = upload_arr(&p->pty, ty_dev, (size_t)n, cudaMemcpyDeviceToDevice);
    if (rc) return rc;
    p->pts_lo = 0; p->pts_hi = p->px;
    if (n > 0) {
        std::vector<double> h((size_t)n);
        CUDA_TRY(cudaMemcpy(h.data(), tx_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
        double lo = h[0], hi = h[0];
        for (int64_t i = 1; i < n; ++i) { lo = std::min(lo, h[i]); hi = std::max(hi, h[i]); }
        p->pts_lo = std::max(0, (int)floor((lo - p->x0) / p->h) - p->m - 1);
        p->pts_hi = std::min(p->px, (int)ceil((hi - p->x0) / p->h) + p->m + 2);
    }
    return IPDE_OK;
}

int ewald_apply(EwaldPlan* p, const double* q, double* grid_out, double* point_out, int accumulate_points, cudaStream_t st) {
    const bool want_grid = (p->n_og > 0 && grid_out != nullptr);
    const bool want_pts = (p->n_pt > 0 && point_out != nullptr);
    if (!want_grid && !want_pts) return IPDE_OK;
    EwaldTab T;
    T.S = p->tabS; T.A = p->tabA; T.nint = p->nint; T.ndeg = p->ndeg; T.nlog = p->nlog;
    T.inv_dz = (double)p->nint / p->rc2;
    const double two_xi2 = 2.0 * p->xi * p->xi;
    const double ck = (p->kind == 1) ? exp(-p->kappa * p->kappa / (4.0 * p->xi * p->xi)) : 1.0;
    const double wnorm = two_xi2 / 3.14159265358979323846;      // (2 xi^2 / pi)
    if (p->ntiles > 0) {
        KERNEL_TAG("k_ewald_tiles", 16.0 * 256.0 * (double)p->ntiles);
        k_ewald_tiles<<<p->ntiles, dim3(EW_TILE, EW_TILE), 0, st>>>(p->tile_ix, p->tile_iy, p->tile_ptr, p->tile_src, p->sx,
                                                                   p->sy, q, T, p->x0, p->y0, p->h, p->rc2, p->rw2, two_xi2,
                                                                   ck * wnorm, p->px, p->py, p->Sw, p->Lc, p->tile_mask,
                                                                   want_grid ? 1 : 0);
        KERNEL_CHECK();
    }
    GridPlan* G = p->fft;
    const double c2r_scale = 1.0 / ((double)G->nx * (double)G->ny) * 2.0;
    // zero padding: only rows [in_lo, in_hi) of S_w are nonzero, only the target rows of the result are needed
    const int out_lo = want_grid ? p->grid_lo : p->pts_lo, out_hi = want_grid ? p->grid_hi : p->pts_hi;
    const int win1[4] = {p->in_lo, p->in_hi, out_lo, out_hi};
    int rc = fft_r2c_rows_nz(G->row, p->Sw + (size_t)p->in_lo * G->ny, p->spec + (size_t)p->in_lo * G->pitch,
                             p->in_hi - p->in_lo, G->ny, G->pitch, p->in_chi, st);
    if (rc) return rc;
    ColMult MU, MH;
    MU.kind = 3; MU.k2c = 0.0; MU.table = p->multU; MU.fmul = 1; MU.ny = 0; MU.col_offset = 0; MU.deriv = 0;
    MH = MU; MH.table = p->multH;
    // first field: forward columns, symbol, inverse columns in one pass (the plain 2-D spectrum is kept in
    // p->spec when a second field follows); second field: symbol + inverse columns from that copy
    rc = grid_cols_symbol(G, p->spec, p->specw, (want_grid && want_pts) ? p->spec : nullptr, 1, want_grid ? MU : MH, st, win1);
    if (rc) return rc;
    rc = fft_c2r_rows_nz(G->row, p->specw + (size_t)out_lo * G->pitch, p->R + (size_t)out_lo * G->ny, out_hi - out_lo, G->pitch,
                         G->ny, c2r_scale, G->row->d.n + 1, st, want_grid ? p->grid_chi : 0);
    if (rc) return rc;
    if (want_grid) {
        const int gb = (int)std::min<int64_t>(cdiv(p->n_og, 256), (int64_t)ipde_num_sms() * 8);
        k_ewald_grid_out<<<gb, 256, 0, st>>>(p->R, p->Lc, p->og_lin, p->og_loc, p->og_out, p->n_og, grid_out);
        KERNEL_CHECK();
    }
    if (want_pts) {
        if (want_grid) {
            const int win2[2] = {p->pts_lo, p->pts_hi};
            rc = grid_cols_symbol_inverse(G, p->spec, p->specw, MH, st, win2);
            if (!rc) rc = fft_c2r_rows(G->row, p->specw + (size_t)p->pts_lo * G->pitch, p->R + (size_t)p->pts_lo * G->ny,
                                       p->pts_hi - p->pts_lo, G->pitch, G->ny, c2r_scale, st);
            if (rc) return rc;
        }
        const int pb = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(p->n_pt * 32, 256), (int64_t)ipde_num_sms() * 16));
        KERNEL_TAG("k_ewald_points", 8.0 * 961.0 * (double)p->n_pt);
        k_ewald_points<<<pb, 256, 0, st>>>(p->R, p->ptx, p->pty, p->n_pt, p->tile_map, p->ntx, p->nty, p->tile_ptr,
                                           p->tile_src, p->sx, p->sy, q, T, p->x0, p->y0, p->h, p->rc2, two_xi2,
                                           wnorm * p->h * p->h, p->m, p->px, p->py, point_out, accumulate_points);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Value AND gradient of the potential, accumulated with affine weights: the building block of the stokeslet
// evaluation (ipde_b200/ewald.py: StokesEvaluator).  Field c (0 value, 1 d/dx, 2 d/dy) of the potential of the
// charges q is added to up to two destination arrays as
//      dst[pos] += (w0 + wx (x - xc) + wy (y - yc)) * field_c(x, y)
// at the grid targets (dst indexed by the targets' out_pos) and at the point targets (dst indexed by point).
struct EwaldTabG {
    const double *S, *A, *dS, *dA;
    int nint, ndeg, nlog;
    double inv_dz;
};

__device__ __forceinline__ double ew_clenshaw(const double* __restrict__ c, int ndeg, double t) {
    const double t2 = 2.0 * t;
    double b0 = 0.0, b1 = 0.0;
    for (int k = ndeg - 1; k >= 1; --k) {
        const double tmp = fma(t2, b0, c[k]) - b1;
        b1 = b0; b0 = tmp;
    }
    return fma(t, b0, c[0]) - b1;
}

// G_loc and dG_loc/dz at z = r^2 (tables in shared memory: S, A, dS, dA consecutively, each nint*ndeg)
__device__ __forceinline__ void ewald_gloc_grad(double z, const double* __restrict__ sh, int nint, int ndeg, int nlog,
                                                double inv_dz, double* g, double* dg) {
    const double u = z * inv_dz;
    int i = (int)u;
    if (i > nint - 1) i = nint - 1;
    const double t = 2.0 * (u - (double)i) - 1.0;
    const int ntab = nint * ndeg;
    double s = ew_clenshaw(sh + i * ndeg, ndeg, t);
    double ds = ew_clenshaw(sh + 2 * ntab + i * ndeg, ndeg, t);
    if (i < nlog) {
        const double a = ew_clenshaw(sh + ntab + i * ndeg, ndeg, t);
        const double da = ew_clenshaw(sh + 3 * ntab + i * ndeg, ndeg, t);
        const double lg = log(z);
        s -= a * lg;
        ds -= da * lg + a / z;
    }
    *g = s; *dg = ds;
}

__global__ void __launch_bounds__(EW_TILE* EW_TILE)
k_ewald_tiles_grad(const int* __restrict__ tile_ix, const int* __restrict__ tile_iy, const int* __restrict__ tile_ptr,
                   const int* __restrict__ tile_src, const double* __restrict__ sx, const double* __restrict__ sy,
                   const double* __restrict__ q, EwaldTabG T, double x0, double y0, double h, double rc2, double rw2,
                   double two_xi2, double norm, int px, int py, double* __restrict__ Sw, double* __restrict__ Lc,
                   double* __restrict__ Lcx, double* __restrict__ Lcy, const unsigned* __restrict__ tile_mask) {
    __shared__ double shx[EW_SCHUNK], shy[EW_SCHUNK], shq[EW_SCHUNK];
    __shared__ double shex[EW_SCHUNK][EW_TILE], shey[EW_SCHUNK][EW_TILE];
    __shared__ double stab[4 * EW_MAX_TAB / 2];      // S, A, dS, dA (nint*ndeg <= 512 each)
    const int tid = threadIdx.y * EW_TILE + threadIdx.x;
    const int ntab = T.nint * T.ndeg;
    for (int i = tid; i < ntab; i += EW_TILE * EW_TILE) {
        stab[i] = T.S[i]; stab[ntab + i] = T.A[i]; stab[2 * ntab + i] = T.dS[i]; stab[3 * ntab + i] = T.dA[i];
    }
    const int t = blockIdx.x;
    const int ix = tile_ix[t] * EW_TILE + threadIdx.y;
    const int iy = tile_iy[t] * EW_TILE + threadIdx.x;
    const double x = x0 + ix * h, y = y0 + iy * h;
    const int s0 = tile_ptr[t], s1 = tile_ptr[t + 1];
    const bool want_local = tile_mask ? ((tile_mask[t * 8 + (tid >> 5)] >> (tid & 31)) & 1u) : false;
    double spread = 0.0, loc = 0.0, lx = 0.0, ly = 0.0;
    for (int base = s0; base < s1; base += EW_SCHUNK) {
        const int cnt = min(EW_SCHUNK, s1 - base);
        __syncthreads();
        if (tid < cnt) {
            const int s = tile_src[base + tid];
            shx[tid] = sx[s]; shy[tid] = sy[s]; shq[tid] = q[s];
        }
        __syncthreads();
        ewald_window_factors(shx, shy, cnt, x0 + tile_ix[t] * EW_TILE * h, y0 + tile_iy[t] * EW_TILE * h, h, rw2, two_xi2,
                             shex, shey, tid);
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            spread = fma(shq[i] * shex[i][threadIdx.y], shey[i][threadIdx.x], spread);
            if (want_local) {
                const double dx = x - shx[i], dy = y - shy[i];
                const double d2 = dx * dx + dy * dy;
                if (d2 < rc2 && d2 > 0.0) {
                    double g, dg;
                    ewald_gloc_grad(d2, stab, T.nint, T.ndeg, T.nlog, T.inv_dz, &g, &dg);
                    loc = fma(shq[i], g, loc);
                    const double c = 2.0 * shq[i] * dg;
                    lx = fma(c, dx, lx); ly = fma(c, dy, ly);
                }
            }
        }
    }
    if (ix < px && iy < py) Sw[(int64_t)ix * py + iy] = spread * norm;
    if (want_local) {
        const int64_t o = (int64_t)t * (EW_TILE * EW_TILE) + tid;
        Lc[o] = loc; Lcx[o] = lx; Lcy[o] = ly;
    }
}

// dst[pos[i]] += (w0 + wx (x - xc) + wy (y - yc)) (U[node] + local[node])
__global__ void k_ewald_grid_out_w(const double* __restrict__ U, const double* __restrict__ Lc, const int64_t* __restrict__ lin,
                                   const int64_t* __restrict__ loc, const int64_t* __restrict__ pos, int64_t n, int py,
                                   double x0, double y0, double h, double xc, double yc, double w0, double wx, double wy,
                                   double* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t li = lin[i];
        double v = U[li];
        const int64_t l = loc[i];
        if (l >= 0) v += Lc[l];
        const double x = x0 + (double)(li / py) * h - xc, y = y0 + (double)(li % py) * h - yc;
        dst[pos[i]] += (w0 + wx * x + wy * y) * v;
    }
}

// one warp per point target: value and gradient (far: Gaussian window and its derivative on H; local: tables)
__global__ void k_ewald_points_grad(const double* __restrict__ H, const double* __restrict__ tx, const double* __restrict__ ty,
                                    int64_t npts, const int* __restrict__ tile_map, int ntx, int nty,
                                    const int* __restrict__ tile_ptr, const int* __restrict__ tile_src,
                                    const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ q,
                                    EwaldTabG T, double x0, double y0, double h, double rc2, double two_xi2, double wnorm, int m,
                                    int px, int py, double* __restrict__ pv) {
    __shared__ double stab[4 * EW_MAX_TAB / 2];
    const int ntab = T.nint * T.ndeg;
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) {
        stab[i] = T.S[i]; stab[ntab + i] = T.A[i]; stab[2 * ntab + i] = T.dS[i]; stab[3 * ntab + i] = T.dA[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int W = 2 * m + 1;
    for (int64_t p = warp; p < npts; p += nwarps) {
        const double x = tx[p], y = ty[p];
        const double X = (x - x0) / h, Y = (y - y0) / h;
        const int i0 = (int)ceil(X - m), j0 = (int)ceil(Y - m);
        // lane b: weights of column b and (same index) of row b, and their derivatives
        double wy = 0.0, wx = 0.0, dwy = 0.0, dwx = 0.0;
        if (lane < W) {
            const double ddx = x - (x0 + (double)(i0 + lane) * h), ddy = y - (y0 + (double)(j0 + lane) * h);
            wx = exp(-two_xi2 * ddx * ddx);
            wy = exp(-two_xi2 * ddy * ddy);
            dwx = -2.0 * two_xi2 * ddx * wx;         // d/dx exp(-2 xi^2 (x - xl)^2)
            dwy = -2.0 * two_xi2 * ddy * wy;
        }
        double a0 = 0.0, a1 = 0.0;     // sum_a wx_a H[a][b], sum_a dwx_a H[a][b]  for this lane's column b
        const bool in_cols = (lane < W) && (j0 + lane >= 0) && (j0 + lane < py);
        for (int a = 0; a < W; ++a) {
            const double wa = __shfl_sync(0xffffffffu, wx, a);
            const double dwa = __shfl_sync(0xffffffffu, dwx, a);
            const int ii = i0 + a;
            if (in_cols && ii >= 0 && ii < px) {
                const double hv = H[(int64_t)ii * py + j0 + lane];
                a0 = fma(wa, hv, a0); a1 = fma(dwa, hv, a1);
            }
        }
        double val = a0 * wy * wnorm, gx = a1 * wy * wnorm, gy = a0 * dwy * wnorm;
        const int fxi = (int)floor(X), fyi = (int)floor(Y);
        const int tix = fxi / EW_TILE, tiy = fyi / EW_TILE;
        if (fxi >= 0 && tix < ntx && fyi >= 0 && tiy < nty) {
            const int t = tile_map[tix * nty + tiy];
            if (t >= 0) {
                const int s0 = tile_ptr[t], s1 = tile_ptr[t + 1];
                for (int s = s0 + lane; s < s1; s += 32) {
                    const int k = tile_src[s];
                    const double dx = x - sx[k], dy = y - sy[k];
                    const double d2 = dx * dx + dy * dy;
                    if (d2 < rc2 && d2 > 0.0) {
                        double g, dg;
                        ewald_gloc_grad(d2, stab, T.nint, T.ndeg, T.nlog, T.inv_dz, &g, &dg);
                        val = fma(q[k], g, val);
                        const double c = 2.0 * q[k] * dg;
                        gx = fma(c, dx, gx); gy = fma(c, dy, gy);
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            val += __shfl_xor_sync(0xffffffffu, val, o);
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
        }
        if (lane == 0) { pv[p] = val; pv[npts + p] = gx; pv[2 * npts + p] = gy; }
    }
}

__global__ void k_ewald_points_accum(const double* __restrict__ v, const double* __restrict__ tx, const double* __restrict__ ty,
                                     int64_t n, double xc, double yc, double w0, double wx, double wy, double* __restrict__ dst) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] += (w0 + wx * (tx[i] - xc) + wy * (ty[i] - yc)) * v[i];
}

int ewald_set_gradient_tables(EwaldPlan* p, const double* dS_host, const double* dA_host) {
    cudaFree(p->tabdS); cudaFree(p->tabdA);
    p->tabdS = p->tabdA = nullptr;
    ARG_CHECK(p->nint * p->ndeg <= EW_MAX_TAB / 2, "local-kernel table too large for the gradient kernels");
    int rc = upload_arr(&p->tabdS, dS_host, (size_t)p->nint * p->ndeg, cudaMemcpyHostToDevice);
    if (!rc) rc = upload_arr(&p->tabdA, dA_host, (size_t)p->nint * p->ndeg, cudaMemcpyHostToDevice);
    return rc;
}

// gdst/pdst: host arrays of 6 device pointers [field][2] (null = unused); gw/pw: host arrays [field][2][3] of
// (w0, wx, wy).  (xc, yc): origin of the affine weights.
int ewald_apply_fields(EwaldPlan* p, const double* q, double xc, double yc, const double* gw, double* const* gdst,
                       const double* pw, double* const* pdst, cudaStream_t st) {
    ARG_CHECK(p->tabdS && p->tabdA, "gradient tables not set (ipde_ewald_set_gradient_tables)");
    bool gf[3] = {false, false, false}, pf[3] = {false, false, false};
    bool want_grid = false, want_pts = false;
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 2; ++d) {
            if (p->n_og > 0 && gdst && gdst[c * 2 + d]) { gf[c] = true; want_grid = true; }
            if (p->n_pt > 0 && pdst && pdst[c * 2 + d]) { pf[c] = true; want_pts = true; }
        }
    if (!want_grid && !want_pts) return IPDE_OK;
    const size_t nloc = (size_t)std::max(1, p->ntiles) * EW_TILE * EW_TILE;
    if (!p->Lcx) {
        CUDA_TRY(cudaMalloc(&p->Lcx, sizeof(double) * nloc));
        CUDA_TRY(cudaMalloc(&p->Lcy, sizeof(double) * nloc));
    }
    if (want_pts && !p->pv) CUDA_TRY(cudaMalloc(&p->pv, sizeof(double) * 3 * (size_t)p->n_pt));
    EwaldTabG T;
    T.S = p->tabS; T.A = p->tabA; T.dS = p->tabdS; T.dA = p->tabdA;
    T.nint = p->nint; T.ndeg = p->ndeg; T.nlog = p->nlog;
    T.inv_dz = (double)p->nint / p->rc2;
    const double two_xi2 = 2.0 * p->xi * p->xi;
    const double ck = (p->kind == 1) ? exp(-p->kappa * p->kappa / (4.0 * p->xi * p->xi)) : 1.0;
    const double wnorm = two_xi2 / 3.14159265358979323846;
    if (p->ntiles > 0) {
        KERNEL_TAG("k_ewald_tiles_grad", 32.0 * 256.0 * (double)p->ntiles);
        k_ewald_tiles_grad<<<p->ntiles, dim3(EW_TILE, EW_TILE), 0, st>>>(p->tile_ix, p->tile_iy, p->tile_ptr, p->tile_src, p->sx,
                                                                        p->sy, q, T, p->x0, p->y0, p->h, p->rc2, p->rw2,
                                                                        two_xi2, ck * wnorm, p->px, p->py, p->Sw, p->Lc, p->Lcx,
                                                                        p->Lcy, want_grid ? p->tile_mask : nullptr);
        KERNEL_CHECK();
    }
    GridPlan* G = p->fft;
    const double c2r_scale = 1.0 / ((double)G->nx * (double)G->ny) * 2.0;
    int rc = fft_r2c_rows(G->row, p->Sw + (size_t)p->in_lo * G->ny, p->spec + (size_t)p->in_lo * G->pitch, p->in_hi - p->in_lo,
                          G->ny, G->pitch, st);
    if (rc) return rc;
    ColMult M0;
    M0.kind = 3; M0.k2c = 0.0; M0.table = p->multU; M0.fmul = 1; M0.ny = 0; M0.col_offset = 0; M0.deriv = 0;
    // The first pass does the forward column transform and leaves the plain 2-D spectrum in p->spec; every further
    // field is one "symbol + inverse columns" pass from that copy followed by the inverse row transform.
    bool first = true;
    const int gb = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(p->n_og, 256), (int64_t)ipde_num_sms() * 8));
    const double* Ls[3] = {p->Lc, p->Lcx, p->Lcy};
    auto field_pass = [&](const double* table, int deriv, int lo, int hi) -> int {
        ColMult M = M0;
        M.table = table; M.deriv = deriv;
        int r;
        if (first) {
            const int win1[4] = {p->in_lo, p->in_hi, lo, hi};
            r = grid_cols_symbol(G, p->spec, p->specw, p->spec, 1, M, st, win1);
            first = false;
        } else {
            const int win2[2] = {lo, hi};
            r = grid_cols_symbol_inverse(G, p->spec, p->specw, M, st, win2);
        }
        if (r) return r;
        return fft_c2r_rows(G->row, p->specw + (size_t)lo * G->pitch, p->R + (size_t)lo * G->ny, hi - lo, G->pitch, G->ny,
                            c2r_scale, st);
    };
    if (want_grid) {
        for (int c = 0; c < 3; ++c) {
            if (!gf[c]) continue;
            rc = field_pass(p->multU, c, p->grid_lo, p->grid_hi);
            if (rc) return rc;
            for (int d = 0; d < 2; ++d) {
                double* dst = gdst[c * 2 + d];
                if (!dst) continue;
                const double* w = gw + (c * 2 + d) * 3;
                k_ewald_grid_out_w<<<gb, 256, 0, st>>>(p->R, Ls[c], p->og_lin, p->og_loc, p->og_out, p->n_og, p->py, p->x0, p->y0,
                                                       p->h, xc, yc, w[0], w[1], w[2], dst);
                KERNEL_CHECK();
            }
        }
    }
    if (want_pts) {
        rc = field_pass(p->multH, 0, p->pts_lo, p->pts_hi);
        if (rc) return rc;
        const int pb = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(p->n_pt * 32, 256), (int64_t)ipde_num_sms() * 16));
        KERNEL_TAG("k_ewald_points_grad", 8.0 * 961.0 * (double)p->n_pt);
        k_ewald_points_grad<<<pb, 256, 0, st>>>(p->R, p->ptx, p->pty, p->n_pt, p->tile_map, p->ntx, p->nty, p->tile_ptr,
                                                p->tile_src, p->sx, p->sy, q, T, p->x0, p->y0, p->h, p->rc2, two_xi2,
                                                wnorm * p->h * p->h, p->m, p->px, p->py, p->pv);
        KERNEL_CHECK();
        const int ab = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(p->n_pt, 256), 2048));
        for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 2; ++d) {
                double* dst = pdst[c * 2 + d];
                if (!dst) continue;
                const double* w = pw + (c * 2 + d) * 3;
                k_ewald_points_accum<<<ab, 256, 0, st>>>(p->pv + (size_t)c * p->n_pt, p->ptx, p->pty, p->n_pt, xc, yc, w[0], w[1],
                                                         w[2], dst);
                KERNEL_CHECK();
            }
    }
    return IPDE_OK;
}

// test/debug access to the work arrays: 0 S_w (px*py), 1 last inverse transform (px*py), 2 local sums (ntiles*256)
int ewald_debug_buffer(EwaldPlan* p, int which, double** ptr, int64_t* n) {
    if (which == 0) { *ptr = p->Sw; *n = (int64_t)p->px * p->py; }
    else if (which == 1) { *ptr = p->R; *n = (int64_t)p->px * p->py; }
    else if (which == 2) { *ptr = p->Lc; *n = (int64_t)p->ntiles * EW_TILE * EW_TILE; }
    else { ipde_set_error("ewald_debug_buffer: unknown buffer %d", which); return IPDE_ERR_ARG; }
    return IPDE_OK;
}
