// Type-2 nonuniform FFT (uniform spectrum -> values at nonuniform points) in FP64 to ~1e-14:
// deconvolved zero-padded spectrum -> sigma-oversampled inverse FFT (our own Stockham) -> "exponential
// of semicircle" kernel interpolation, one warp per target.
//
// Replaces near_finder.nufft_interp2d.periodic_interp2d / finufft.nufft2d2(isign=+1, eps=1e-14,
// modeord=1) at ipde/solvers/multi_boundary/scalar.py:80-88 and ipde/ebdy_collection.py:604-614:
//      v_j = Re (1/(nx ny)) sum_k fh[k] (i kx)^dx (i ky)^dy exp(+i (k1 x_j + k2 y_j))
// The unpaired Nyquist modes (k = -n/2 in numpy order) are handled so that the real part of the
// reference's complex sum is reproduced exactly ("Nyquist modes").
#include "nufft.cuh"
#include <algorithm>

static const double TWO_PI = 6.283185307179586476925286766559;

int nufft2_plan_create(int nx, int ny, double xh, double yh, int fx, int fy, int w, double beta, const double* px_host,
                       const double* py_host, Nufft2Plan** out) {
    ARG_CHECK(w >= 2 && w <= 32, "kernel width must be in [2, 32]");
    ARG_CHECK(fx >= nx + w && fy >= ny + w, "fine grid too small");
    Nufft2Plan* p = new Nufft2Plan();
    p->nx = nx; p->ny = ny; p->fx = fx; p->fy = fy; p->w = w; p->beta = beta;
    p->xh = xh; p->yh = yh;
    // fine grid spacing only matters through the wavenumbers, which the pad kernel takes from the coarse grid
    int rc = grid_plan_create(fx, fy, 1.0, 1.0, &p->fine);
    if (rc) { delete p; return rc; }
    CUDA_TRY(cudaMalloc(&p->px, sizeof(double) * nx));
    CUDA_TRY(cudaMalloc(&p->py, sizeof(double) * (ny / 2 + 1)));
    CUDA_TRY(cudaMemcpy(p->px, px_host, sizeof(double) * nx, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(p->py, py_host, sizeof(double) * (ny / 2 + 1), cudaMemcpyHostToDevice));
    const size_t nspec = (size_t)fx * (fy / 2 + 1);
    CUDA_TRY(cudaMalloc(&p->fine_spec, sizeof(cplx) * nspec));
    CUDA_TRY(cudaMemset(p->fine_spec, 0, sizeof(cplx) * nspec));     // columns beyond ny/2 stay zero for good
    CUDA_TRY(cudaMalloc(&p->fine_real, sizeof(double) * (size_t)fx * fy));
    *out = p;
    return IPDE_OK;
}

void nufft2_plan_destroy(Nufft2Plan* p) {
    if (!p) return;
    grid_plan_destroy(p->fine);
    cudaFree(p->px); cudaFree(p->py); cudaFree(p->fine_spec); cudaFree(p->fine_real); cudaFree(p->mask);
    delete p;
}

__device__ __forceinline__ int row_of_freq(int i, int n1, int n2) {
    if (n2 == 1) return i;
    return n2 * (i % n1) + i / n1;
}
__device__ __forceinline__ int freq_of_row2(int pos, int n1, int n2) {
    if (n2 == 1) return pos;
    return pos / n2 + n1 * (pos % n2);
}

// Build the padded, deconvolved fine half-spectrum from a coarse half-spectrum.
__global__ void k_nufft_pad(const cplx* __restrict__ spec, int nx, int ny, int64_t cpitch, int cn1, int cn2,
                            cplx* __restrict__ fine, int fx, int fy, int64_t fpitch, int fn1, int fn2,
                            const double* __restrict__ px, const double* __restrict__ py, double xh, double yh, int dxp,
                            int dyp, double scale) {
    // only the columns j <= ny/2 can be nonzero; the rest of the fine spectrum was zeroed at plan creation and is
    // never written (the column passes are restricted to the same columns)
    const int64_t ncols = ny / 2 + 1;
    const int64_t total = (int64_t)fx * ncols;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int pos = (int)(t / ncols), j = (int)(t % ncols);
        const int I = freq_of_row2(pos, fn1, fn2);
        const int sI = (I < (fx + 1) / 2) ? I : I - fx;      // signed fine row frequency
        cplx v = make_double2(0.0, 0.0);
        const int hx = nx / 2, hy = ny / 2;
        const bool x_in = (2 * abs(sI) < nx) || (nx % 2 == 0 && abs(sI) == hx);
        if (x_in && j <= hy) {
            const bool x_nyq = (nx % 2 == 0 && abs(sI) == hx);
            const bool y_nyq = (j == hy);
            double wgt = 1.0;
            if (x_nyq) wgt *= 0.5;
            if (y_nyq) wgt *= 0.5;
            if (x_nyq && y_nyq) wgt = (sI > 0) ? 0.5 : 0.0;   // exp(-i(nx/2 x + ny/2 y)) pairs only with (+,+)
            const int ci = (sI >= 0) ? sI : sI + nx;          // coarse frequency index (nx/2 for both +-nx/2)
            cplx a = spec[(int64_t)row_of_freq(ci, cn1, cn2) * cpitch + j];
            const double kx = TWO_PI * (double)sI / ((double)nx * xh);
            const double ky = TWO_PI * (double)j / ((double)ny * yh);
            a = cscale(a, wgt * scale * px[ci] * py[j]);
            if (dxp) a = cscale(cmuli(a), kx);
            if (dyp) a = cscale(cmuli(a), ky);
            v = a;
        }
        fine[(int64_t)pos * fpitch + j] = v;
    }
}

// One warp per target; lanes split the w*w stencil.
__global__ void k_nufft_interp(const double* __restrict__ fine, int fx, int fy, const double* __restrict__ x,
                               const double* __restrict__ y, int64_t npts, int w, double beta, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double half = 0.5 * w;
    for (int64_t p = warp; p < npts; p += nwarps) {
        const double X = x[p] * ((double)fx / TWO_PI), Y = y[p] * ((double)fy / TWO_PI);
        const int i0 = (int)ceil(X - half), j0 = (int)ceil(Y - half);
        double acc = 0.0;
        const int npt = w * w;
        for (int s = lane; s < npt; s += 32) {
            const int a = s / w, b = s % w;
            const double zx = ((double)(i0 + a) - X) / half, zy = ((double)(j0 + b) - Y) / half;
            const double ax = 1.0 - zx * zx, ay = 1.0 - zy * zy;
            if (ax >= 0.0 && ay >= 0.0) {
                const double k = exp(beta * (sqrt(ax) - 1.0) + beta * (sqrt(ay) - 1.0));
                int ii = (i0 + a) % fx; if (ii < 0) ii += fx;
                int jj = (j0 + b) % fy; if (jj < 0) jj += fy;
                acc += k * fine[(int64_t)ii * fy + jj];
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[p] = acc;
    }
}

// value and both first derivatives from ONE oversampled grid: the derivatives come from differentiating the
// interpolation kernel (d/dx phi(x - x_l)), which is as accurate as multiplying the spectrum by i k (measured:
// 5e-14 of max|grad| on a full-band random spectrum) and saves two of the three fine-grid transforms.
__global__ void k_nufft_interp3(const double* __restrict__ fine, int fx, int fy, const double* __restrict__ x,
                                const double* __restrict__ y, int64_t npts, int w, double beta, double gx, double gy,
                                double* __restrict__ out_v, double* __restrict__ out_dx, double* __restrict__ out_dy) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double half = 0.5 * w;
    const double sx = (double)fx / TWO_PI, sy = (double)fy / TWO_PI;
    for (int64_t p = warp; p < npts; p += nwarps) {
        const double X = x[p] * sx, Y = y[p] * sy;
        const int i0 = (int)ceil(X - half), j0 = (int)ceil(Y - half);
        double av = 0.0, adx = 0.0, ady = 0.0;
        const int npt = w * w;
        for (int s = lane; s < npt; s += 32) {
            const int a = s / w, b = s % w;
            const double zx = ((double)(i0 + a) - X) / half, zy = ((double)(j0 + b) - Y) / half;
            const double ax = 1.0 - zx * zx, ay = 1.0 - zy * zy;
            if (ax > 0.0 && ay > 0.0) {
                const double rx = sqrt(ax), ry = sqrt(ay);
                const double k = exp(beta * (rx - 1.0) + beta * (ry - 1.0));
                int ii = (i0 + a) % fx; if (ii < 0) ii += fx;
                int jj = (j0 + b) % fy; if (jj < 0) jj += fy;
                const double g = k * fine[(int64_t)ii * fy + jj];
                av += g;
                // d/dx: z = (l - x s)/half, dz/dx = -s/half; d phi/dz = -beta z / sqrt(1 - z^2) phi.  The derivative of
                // the kernel is singular at the edge of its support (phi ~ e^{-beta} there): for a target within ~1e-10
                // of a fine-grid node the edge node has r -> 0 and the factor 1/r amplifies round-off to 1e-11; below
                // r = 0.02 the term is at most 3e-14 of the gradient scale and is dropped
                if (rx > 0.02) adx += g * (beta * zx / rx) * (sx / half);
                if (ry > 0.02) ady += g * (beta * zy / ry) * (sy / half);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            av += __shfl_xor_sync(0xffffffffu, av, o);
            adx += __shfl_xor_sync(0xffffffffu, adx, o);
            ady += __shfl_xor_sync(0xffffffffu, ady, o);
        }
        if (lane == 0) { out_v[p] = av; out_dx[p] = adx * gx; out_dy[p] = ady * gy; }
    }
}

// the fine grid is only read on the w x w stencils of the targets: one bit per (fine row, block of 64 reals)
__global__ void k_nufft_mark(int fx, int fy, const double* __restrict__ x, const double* __restrict__ y, int64_t npts, int w,
                             uint32_t* __restrict__ mask, int words) {
    const double half = 0.5 * w;
    const int64_t total = npts * w;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = t / w;
        const int a = (int)(t - p * w);
        const double X = x[p] * ((double)fx / TWO_PI), Y = y[p] * ((double)fy / TWO_PI);
        const int i0 = (int)ceil(X - half), j0 = (int)ceil(Y - half);
        int ii = (i0 + a) % fx; if (ii < 0) ii += fx;
        int ja = j0 % fy; if (ja < 0) ja += fy;
        int jb = (j0 + w - 1) % fy; if (jb < 0) jb += fy;
        uint32_t* mrow = mask + (int64_t)ii * words;
        atomicOr(&mrow[ja >> 11], 1u << ((ja >> 6) & 31));
        atomicOr(&mrow[jb >> 11], 1u << ((jb >> 6) & 31));      // w <= 32: the stencil touches at most two blocks
    }
}

// inverse transform of the zero-padded fine spectrum: column passes over the nonzero columns only; the row pass
// stores only the blocks of the fine grid that the targets' stencils read
static int nufft2_inverse(Nufft2Plan* p, const double* x, const double* y, int64_t npts, cudaStream_t st) {
    GridPlan* f = p->fine;
    int rc = grid_cols_inverse_n(f, p->fine_spec, p->fine_spec, p->ny / 2 + 1, st);
    if (rc) return rc;
    const uint32_t* mask = nullptr;
    const int words = f->ny / 2048;
    if (npts > 0 && f->ny % 2048 == 0 && p->w <= 32) {
        if (!p->mask) CUDA_TRY(cudaMalloc(&p->mask, sizeof(uint32_t) * (size_t)f->nx * words));
        CUDA_TRY(cudaMemsetAsync(p->mask, 0, sizeof(uint32_t) * (size_t)f->nx * words, st));
        const int mb = (int)std::min<int64_t>(cdiv(npts * p->w, 256), (int64_t)ipde_num_sms() * 16);
        k_nufft_mark<<<mb, 256, 0, st>>>(f->nx, f->ny, x, y, npts, p->w, p->mask, words);
        KERNEL_CHECK();
        mask = p->mask;
    }
    return fft_c2r_rows_nz(f->row, p->fine_spec, p->fine_real, f->nx, f->pitch, f->ny,
                           1.0 / ((double)f->nx * (double)f->ny) * 2.0, p->ny / 2 + 1, st, 0, mask, words);
}

int nufft2_apply3(Nufft2Plan* p, const cplx* spec, int cn1, int cn2, const double* x, const double* y, int64_t npts,
                  double* out_v, double* out_dx, double* out_dy, cudaStream_t st) {
    GridPlan* f = p->fine;
    const int64_t total = (int64_t)p->fx * (p->ny / 2 + 1);
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)ipde_num_sms() * 8);
    const double scale = ((double)p->fx * (double)p->fy) / ((double)p->nx * (double)p->ny);
    k_nufft_pad<<<blocks, 256, 0, st>>>(spec, p->nx, p->ny, p->ny / 2 + 1, cn1, cn2, p->fine_spec, p->fx, p->fy, f->pitch,
                                        f->n1, f->n2, p->px, p->py, p->xh, p->yh, 0, 0, scale);
    KERNEL_CHECK();
    int rc = nufft2_inverse(p, x, y, npts, st);
    if (rc) return rc;
    if (npts > 0) {
        const int b2 = (int)std::min<int64_t>(cdiv(npts * 32, 256), (int64_t)ipde_num_sms() * 16);
        k_nufft_interp3<<<b2, 256, 0, st>>>(p->fine_real, p->fx, p->fy, x, y, npts, p->w, p->beta,
                                            TWO_PI / ((double)p->nx * p->xh), TWO_PI / ((double)p->ny * p->yh), out_v, out_dx, out_dy);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}

int nufft2_apply(Nufft2Plan* p, const cplx* spec, int cn1, int cn2, int dxp, int dyp, const double* x, const double* y,
                 int64_t npts, double* out, cudaStream_t st) {
    GridPlan* f = p->fine;
    const int64_t total = (int64_t)p->fx * (p->ny / 2 + 1);
    const int blocks = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)ipde_num_sms() * 8);
    const double scale = ((double)p->fx * (double)p->fy) / ((double)p->nx * (double)p->ny);
    k_nufft_pad<<<blocks, 256, 0, st>>>(spec, p->nx, p->ny, p->ny / 2 + 1, cn1, cn2, p->fine_spec, p->fx, p->fy, f->pitch,
                                        f->n1, f->n2, p->px, p->py, p->xh, p->yh, dxp, dyp, scale);
    KERNEL_CHECK();
    int rc = nufft2_inverse(p, x, y, npts, st);
    if (rc) return rc;
    if (npts > 0) {
        const int b2 = (int)std::min<int64_t>(cdiv(npts * 32, 256), (int64_t)ipde_num_sms() * 16);
        k_nufft_interp<<<b2, 256, 0, st>>>(p->fine_real, p->fx, p->fy, x, y, npts, p->w, p->beta, out);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// 1-D variant used for radial -> points: rows of a (nrows, n) spectrum, targets carry a row-independent t.
int nufft1_plan_create(int n, int fn, int w, double beta, const double* p_host, Nufft1Plan** out) {
    ARG_CHECK(w >= 2 && w <= 32, "kernel width must be in [2, 32]");
    Nufft1Plan* p = new Nufft1Plan();
    p->n = n; p->fn = fn; p->w = w; p->beta = beta;
    int rc = fft_plan_create(n, &p->coarse);
    if (!rc) rc = fft_plan_create(fn, &p->fine);
    if (rc) { nufft1_plan_destroy(p); return rc; }
    CUDA_TRY(cudaMalloc(&p->dec, sizeof(double) * n));
    CUDA_TRY(cudaMemcpy(p->dec, p_host, sizeof(double) * n, cudaMemcpyHostToDevice));
    *out = p;
    return IPDE_OK;
}

void nufft1_plan_destroy(Nufft1Plan* p) {
    if (!p) return;
    fft_plan_destroy(p->coarse);
    fft_plan_destroy(p->fine);
    cudaFree(p->dec); cudaFree(p->buf_c); cudaFree(p->buf_f);
    delete p;
}

__global__ void k_real_to_cplx(const double* __restrict__ in, cplx* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = make_double2(in[i], 0.0);
}

// pad a (rows, n) spectrum to (rows, fn) with deconvolution, Nyquist split
__global__ void k_nufft1_pad(const cplx* __restrict__ spec, cplx* __restrict__ fine, int rows, int n, int fn,
                             const double* __restrict__ dec, double scale) {
    const int64_t total = (int64_t)rows * fn;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(t / fn), I = (int)(t % fn);
        const int sI = (I < (fn + 1) / 2) ? I : I - fn;
        cplx v = make_double2(0.0, 0.0);
        const bool nyq = (n % 2 == 0 && abs(sI) == n / 2);
        if (2 * abs(sI) < n || nyq) {
            const int ci = (sI >= 0) ? sI : sI + n;
            v = cscale(spec[(int64_t)r * n + ci], (nyq ? 0.5 : 1.0) * scale * dec[ci]);
        }
        fine[t] = v;
    }
}

int nufft1_upsample_rows(Nufft1Plan* p, const double* rows_real, int nrows, cudaStream_t st) {
    const size_t need_c = (size_t)nrows * p->n, need_f = (size_t)nrows * p->fn;
    if (p->cap_rows < nrows) {
        cudaFree(p->buf_c); cudaFree(p->buf_f);
        p->buf_c = nullptr; p->buf_f = nullptr; p->cap_rows = 0;
        CUDA_TRY(cudaMalloc(&p->buf_c, sizeof(cplx) * need_c));
        CUDA_TRY(cudaMalloc(&p->buf_f, sizeof(cplx) * need_f));
        p->cap_rows = nrows;
    }
    const int b1 = (int)std::min<int64_t>(cdiv((int64_t)need_c, 256), (int64_t)ipde_num_sms() * 8);
    k_real_to_cplx<<<b1, 256, 0, st>>>(rows_real, p->buf_c, (int64_t)need_c);
    KERNEL_CHECK();
    int rc = fft_rows(p->coarse, p->buf_c, p->buf_c, nrows, p->n, p->n, -1, 1.0, st);
    if (rc) return rc;
    const int b2 = (int)std::min<int64_t>(cdiv((int64_t)need_f, 256), (int64_t)ipde_num_sms() * 8);
    k_nufft1_pad<<<b2, 256, 0, st>>>(p->buf_c, p->buf_f, nrows, p->n, p->fn, p->dec, 1.0 / p->n);
    KERNEL_CHECK();
    return fft_rows(p->fine, p->buf_f, p->buf_f, nrows, p->fn, p->fn, +1, 1.0, st);
}
