// FP64 Stockham FFT kernels (shared-memory staged, mixed radix 2/3/4/5/7/8/11/13) and a
// global-memory pass fallback for lines that do not fit one CTA's shared memory.
// Replaces the reference's numpy.fft / mkl_fft calls (ipde/utilities.py:5-12).
#include "fft.cuh"
#include <algorithm>

// ------------------------------------------------------------------------------------------
// plans
static bool factorize(int n, std::vector<int>& fac) {
    fac.clear();
    if (n < 1) return false;
    while (n % 8 == 0) { fac.push_back(8); n /= 8; }
    while (n % 4 == 0) { fac.push_back(4); n /= 4; }
    while (n % 2 == 0) { fac.push_back(2); n /= 2; }
    const int odd[] = {3, 5, 7, 11, 13};
    for (int p : odd)
        while (n % p == 0) { fac.push_back(p); n /= p; }
    return n == 1 && (int)fac.size() <= IPDE_FFT_MAX_FACTORS;
}

bool fft_size_supported(int n) {
    std::vector<int> f;
    return factorize(n, f);
}

int fft_next_size(int n) {
    if (n < 2) n = 2;
    if (n % 2) ++n;
    for (;; n += 2) {
        int m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return n;
    }
}

static void fill_twiddles(std::vector<cplx>& t, int64_t n, int64_t count, int64_t denom) {
    t.resize(count);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int64_t j = 0; j < count; ++j) {
        long double a = -tp * (long double)j / (long double)denom;
        t[j] = make_double2((double)cosl(a), (double)sinl(a));
    }
    (void)n;
}

int fft_plan_create(int n, FftPlan** out) {
    std::vector<int> fac;
    if (!factorize(n, fac)) {
        ipde_set_error("fft_plan_create: length %d has a prime factor > 13 (unsupported)", n);
        return IPDE_ERR_UNSUPPORTED;
    }
    FftPlan* p = new FftPlan();
    p->d.n = n;
    p->d.nfac = (int)fac.size();
    for (int i = 0; i < p->d.nfac; ++i) p->d.fac[i] = fac[i];
    std::vector<cplx> tw;
    fill_twiddles(tw, n, n, n);
    CUDA_TRY(cudaMalloc(&p->tw_dev, sizeof(cplx) * (size_t)n));
    CUDA_TRY(cudaMemcpy(p->tw_dev, tw.data(), sizeof(cplx) * (size_t)n, cudaMemcpyHostToDevice));
    fill_twiddles(tw, n, n + 1, 2 * (int64_t)n);
    CUDA_TRY(cudaMalloc(&p->half_tw_dev, sizeof(cplx) * (size_t)(n + 1)));
    CUDA_TRY(cudaMemcpy(p->half_tw_dev, tw.data(), sizeof(cplx) * (size_t)(n + 1), cudaMemcpyHostToDevice));
    p->d.tw = p->tw_dev;
    *out = p;
    return IPDE_OK;
}

void fft_plan_destroy(FftPlan* p) {
    if (!p) return;
    cudaFree(p->tw_dev);
    cudaFree(p->half_tw_dev);
    cudaFree(p->work);
    delete p;
}

static int ensure_work(FftPlan* p, size_t elems) {
    if (p->work_elems >= elems) return IPDE_OK;
    cudaFree(p->work);
    p->work = nullptr;
    p->work_elems = 0;
    CUDA_TRY(cudaMalloc(&p->work, sizeof(cplx) * elems));
    p->work_elems = elems;
    return IPDE_OK;
}

static int pick_threads(int64_t items) {
    if (items >= 2048) return 512;
    if (items >= 512) return 256;
    if (items >= 128) return 128;
    return 64;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// contiguous rows, one line per CTA
__global__ void k_fft_rows(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t in_pitch,
                           int64_t out_pitch, int sign, double scale) {
    extern __shared__ cplx smem[];
    const int n = P.n;
    cplx* a = smem;
    cplx* b = smem + n;
    const cplx* src = in + (int64_t)blockIdx.x * in_pitch;
    cplx* dst = out + (int64_t)blockIdx.x * out_pitch;
    for (int e = threadIdx.x; e < n; e += blockDim.x) a[e] = src[e];
    __syncthreads();
    cplx* r = stockham_smem(a, b, P, 1, sign, threadIdx.x, blockDim.x);
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = cscale(r[e], scale);
}

// strided lines, C adjacent lines per CTA, optional four-step twiddle
__global__ void k_fft_cols(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t elem_stride,
                           int64_t n_inner, int64_t outer_stride, int C, int sign, double scale, int tw_mode,
                           const cplx* __restrict__ tw_big, int64_t n_big) {
    extern __shared__ cplx smem[];
    const int n = P.n;
    cplx* a = smem;
    cplx* b = smem + (size_t)n * C;
    const int64_t tiles = (n_inner + C - 1) / C;
    const int64_t o = blockIdx.x / tiles;
    const int64_t c0 = (blockIdx.x % tiles) * C;
    const int cw = (int)min((int64_t)C, n_inner - c0);
    const int64_t base = o * outer_stride + c0;
    const int total = n * C;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % C, e = i / C;
        cplx v = make_double2(0.0, 0.0);
        if (c < cw) {
            v = in[base + (int64_t)e * elem_stride + c];
            if (tw_mode == 2) {
                cplx w = tw_big[(o * (int64_t)e) % n_big];
                v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
        }
        a[i] = v;
    }
    __syncthreads();
    cplx* r = stockham_smem(a, b, P, C, sign, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i % C, e = i / C;
        if (c < cw) {
            cplx v = r[i];
            if (tw_mode == 1) {
                cplx w = tw_big[(o * (int64_t)e) % n_big];
                v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
            out[base + (int64_t)e * elem_stride + c] = cscale(v, scale);
        }
    }
}

// global-memory Stockham pass (any line length): one work item per butterfly
template <int R>
__global__ void k_fft_pass_global(FftPlanDev P, const cplx* __restrict__ src, cplx* __restrict__ dst, int64_t batch,
                                  int64_t src_line, int64_t src_elem, int64_t dst_line, int64_t dst_elem, int Ns,
                                  int sign, double scale) {
    const int n = P.n;
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const int64_t total = batch * m;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t line = it / m;
        const int j = (int)(it % m);
        const int k = j % Ns;
        const cplx* s = src + line * src_line;
        cplx* d = dst + line * dst_line;
        cplx v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = s[(int64_t)(j + t * m) * src_elem];
        if (Ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) {
                cplx w = P.tw[t * k * tstep];
                v[t] = (sign < 0) ? cmul(v[t], w) : cmulc(v[t], w);
            }
        }
        dft_small<R>(v, sign, P);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) d[(int64_t)(j0 + t * Ns) * dst_elem] = cscale(v[t], scale);
    }
}

// Lines too long for shared memory: pass-per-launch through `work` (contiguous lines of n).
static int fft_lines_global(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_line, int64_t in_elem,
                            int64_t out_line, int64_t out_elem, int sign, double scale, cudaStream_t st) {
    const int n = p->d.n;
    int rc = ensure_work(p, (size_t)2 * n * batch);
    if (rc) return rc;
    cplx* w0 = p->work;
    cplx* w1 = p->work + (size_t)n * batch;
    const cplx* src = in;
    int64_t sl = in_line, se = in_elem;
    int Ns = 1;
    for (int f = 0; f < p->d.nfac; ++f) {
        const int R = p->d.fac[f];
        const bool last = (f == p->d.nfac - 1);
        cplx* dst = last ? out : ((f % 2 == 0) ? w0 : w1);
        int64_t dl = last ? out_line : n, de = last ? out_elem : 1;
        const double sc = last ? scale : 1.0;
        const int64_t total = batch * (n / R);
        const int threads = 256;
        const int blocks = (int)std::min<int64_t>(cdiv(total, threads), (int64_t)ipde_num_sms() * 16);
#define LAUNCH(RR) k_fft_pass_global<RR><<<blocks, threads, 0, st>>>(p->d, src, dst, batch, sl, se, dl, de, Ns, sign, sc)
        switch (R) {
            case 2: LAUNCH(2); break;
            case 3: LAUNCH(3); break;
            case 4: LAUNCH(4); break;
            case 5: LAUNCH(5); break;
            case 7: LAUNCH(7); break;
            case 8: LAUNCH(8); break;
            case 11: LAUNCH(11); break;
            case 13: LAUNCH(13); break;
        }
#undef LAUNCH
        KERNEL_CHECK();
        src = dst; sl = dl; se = de;
        Ns *= R;
    }
    if (p->d.nfac == 0 && out != in) {   // n == 1
        CUDA_TRY(cudaMemcpyAsync(out, in, sizeof(cplx) * batch, cudaMemcpyDeviceToDevice, st));
    }
    return IPDE_OK;
}

int fft_rows(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
             double scale, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (n > IPDE_FFT_SMEM_MAX) return fft_lines_global(p, in, out, batch, in_pitch, 1, out_pitch, 1, sign, scale, st);
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem(k_fft_rows, smem);
    if (rc) return rc;
    k_fft_rows<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, in, out, in_pitch, out_pitch, sign, scale);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_cols(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
             int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
             cudaStream_t st) {
    if (n_inner <= 0 || n_outer <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (n > IPDE_FFT_SMEM_MAX) {
        if (tw_mode != 0) {
            ipde_set_error("fft_cols: twiddled transform of length %d does not fit shared memory", n);
            return IPDE_ERR_UNSUPPORTED;
        }
        // treat every (outer, inner) pair as its own strided line
        for (int64_t o = 0; o < n_outer; ++o) {
            int rc = fft_lines_global(p, in + o * outer_stride, out + o * outer_stride, n_inner, 1, elem_stride, 1,
                                      elem_stride, sign, scale, st);
            if (rc) return rc;
        }
        return IPDE_OK;
    }
    int C = IPDE_FFT_SMEM_MAX / n;
    if (C > 16) C = 16;
    if (C > n_inner) C = (int)n_inner;
    if (C < 1) C = 1;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n * C;
    int rc = set_smem(k_fft_cols, smem);
    if (rc) return rc;
    const int64_t tiles = cdiv(n_inner, C);
    k_fft_cols<<<(unsigned)(tiles * n_outer), pick_threads((int64_t)n * C / 4), smem, st>>>(
        p->d, in, out, elem_stride, n_inner, outer_stride, C, sign, scale, tw_mode, tw_big, n_big);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// real rows
__global__ void k_r2c_rows(FftPlanDev P, const cplx* __restrict__ half_tw, const double* __restrict__ in,
                           cplx* __restrict__ out, int64_t in_pitch, int64_t out_pitch) {
    extern __shared__ cplx smem[];
    const int n = P.n;   // complex length = real length / 2
    cplx* a = smem;
    cplx* b = smem + n;
    const cplx* src = reinterpret_cast<const cplx*>(in + (int64_t)blockIdx.x * in_pitch);
    cplx* dst = out + (int64_t)blockIdx.x * out_pitch;
    for (int e = threadIdx.x; e < n; e += blockDim.x) a[e] = src[e];
    __syncthreads();
    cplx* z = stockham_smem(a, b, P, 1, -1, threadIdx.x, blockDim.x);
    for (int k = threadIdx.x; k <= n; k += blockDim.x) {
        cplx zk = z[k == n ? 0 : k];
        cplx zc = cconj(z[k == 0 ? 0 : n - k]);
        cplx E = cscale(cadd(zk, zc), 0.5);
        cplx O = cmulni(cscale(csub(zk, zc), 0.5));
        dst[k] = cadd(E, cmul(half_tw[k], O));
    }
}

__global__ void k_c2r_rows(FftPlanDev P, const cplx* __restrict__ half_tw, const cplx* __restrict__ in,
                           double* __restrict__ out, int64_t in_pitch, int64_t out_pitch, double scale) {
    extern __shared__ cplx smem[];
    const int n = P.n;
    cplx* a = smem;
    cplx* b = smem + n;
    const cplx* src = in + (int64_t)blockIdx.x * in_pitch;
    cplx* dst = reinterpret_cast<cplx*>(out + (int64_t)blockIdx.x * out_pitch);
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        cplx xk = src[k];
        cplx xc = cconj(src[n - k]);
        cplx E = cscale(cadd(xk, xc), 0.5);
        cplx O = cmulc(cscale(csub(xk, xc), 0.5), half_tw[k]);
        a[k] = cadd(E, cmuli(O));
    }
    __syncthreads();
    cplx* z = stockham_smem(a, b, P, 1, +1, threadIdx.x, blockDim.x);
    for (int e = threadIdx.x; e < n; e += blockDim.x) dst[e] = cscale(z[e], scale);
}

// Global-memory versions of the pack / unpack steps for long rows
__global__ void k_r2c_untangle(int n, const cplx* __restrict__ half_tw, const cplx* __restrict__ z, cplx* __restrict__ out,
                               int64_t batch, int64_t z_pitch, int64_t out_pitch) {
    const int64_t total = batch * (n + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (n + 1);
        const int k = (int)(i % (n + 1));
        const cplx* zz = z + row * z_pitch;
        cplx zk = zz[k == n ? 0 : k];
        cplx zc = cconj(zz[k == 0 ? 0 : n - k]);
        cplx E = cscale(cadd(zk, zc), 0.5);
        cplx O = cmulni(cscale(csub(zk, zc), 0.5));
        out[row * out_pitch + k] = cadd(E, cmul(half_tw[k], O));
    }
}
__global__ void k_c2r_tangle(int n, const cplx* __restrict__ half_tw, const cplx* __restrict__ in, cplx* __restrict__ z,
                             int64_t batch, int64_t in_pitch, int64_t z_pitch) {
    const int64_t total = batch * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / n;
        const int k = (int)(i % n);
        const cplx* src = in + row * in_pitch;
        cplx xk = src[k];
        cplx xc = cconj(src[n - k]);
        cplx E = cscale(cadd(xk, xc), 0.5);
        cplx O = cmulc(cscale(csub(xk, xc), 0.5), half_tw[k]);
        z[row * z_pitch + k] = cadd(E, cmuli(O));
    }
}

int fft_r2c_rows(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (n > IPDE_FFT_SMEM_MAX) {
        // complex FFT of the packed rows into the front of each output row, then untangle in place is not
        // possible (aliasing), so go through the plan's work buffer
        int rc = ensure_work(p, (size_t)3 * n * batch);
        if (rc) return rc;
        cplx* z = p->work + (size_t)2 * n * batch;
        rc = fft_lines_global(p, reinterpret_cast<const cplx*>(in), z, batch, in_pitch / 2, 1, n, 1, -1, 1.0, st);
        if (rc) return rc;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * (n + 1), 256), (int64_t)ipde_num_sms() * 16);
        k_r2c_untangle<<<blocks, 256, 0, st>>>(n, p->half_tw_dev, z, out, batch, n, out_pitch);
        KERNEL_CHECK();
        return IPDE_OK;
    }
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem(k_r2c_rows, smem);
    if (rc) return rc;
    k_r2c_rows<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, p->half_tw_dev, in, out, in_pitch, out_pitch);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_c2r_rows(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 double scale, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (n > IPDE_FFT_SMEM_MAX) {
        int rc = ensure_work(p, (size_t)3 * n * batch);
        if (rc) return rc;
        cplx* z = p->work + (size_t)2 * n * batch;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * n, 256), (int64_t)ipde_num_sms() * 16);
        k_c2r_tangle<<<blocks, 256, 0, st>>>(n, p->half_tw_dev, in, z, batch, in_pitch, n);
        KERNEL_CHECK();
        return fft_lines_global(p, z, reinterpret_cast<cplx*>(out), batch, n, 1, out_pitch / 2, 1, +1, scale, st);
    }
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem(k_c2r_rows, smem);
    if (rc) return rc;
    k_c2r_rows<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, p->half_tw_dev, in, out, in_pitch, out_pitch, scale);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// out = FFT(m * IFFT(in)) / n, fused in shared memory
__global__ void k_fft_mul_fft_rows(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out,
                                   const double* __restrict__ m) {
    extern __shared__ cplx smem[];
    const int n = P.n;
    cplx* a = smem;
    cplx* b = smem + n;
    const int64_t off = (int64_t)blockIdx.x * n;
    for (int e = threadIdx.x; e < n; e += blockDim.x) a[e] = in[off + e];
    __syncthreads();
    cplx* r = stockham_smem(a, b, P, 1, +1, threadIdx.x, blockDim.x);
    const double inv = 1.0 / n;
    for (int e = threadIdx.x; e < n; e += blockDim.x) r[e] = cscale(r[e], m[off + e] * inv);
    __syncthreads();
    cplx* other = (r == a) ? b : a;
    cplx* f = stockham_smem(r, other, P, 1, -1, threadIdx.x, blockDim.x);
    for (int e = threadIdx.x; e < n; e += blockDim.x) out[off + e] = f[e];
}

__global__ void k_scale_by_real(cplx* __restrict__ z, const double* __restrict__ m, int64_t total, double s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = cscale(z[i], m[i] * s);
}

int fft_mul_fft_rows(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (n > IPDE_FFT_SMEM_MAX) {
        int rc = fft_lines_global(p, in, out, batch, n, 1, n, 1, +1, 1.0, st);
        if (rc) return rc;
        const int64_t total = batch * n;
        const int blocks = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)ipde_num_sms() * 16);
        k_scale_by_real<<<blocks, 256, 0, st>>>(out, m, total, 1.0 / n);
        KERNEL_CHECK();
        return fft_lines_global(p, out, out, batch, n, 1, n, 1, -1, 1.0, st);
    }
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem(k_fft_mul_fft_rows, smem);
    if (rc) return rc;
    k_fft_mul_fft_rows<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, in, out, m);
    KERNEL_CHECK();
    return IPDE_OK;
}
