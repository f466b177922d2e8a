// Generic-radix fallback kernels (ping-pong shared-memory Stockham, one line or one column tile per CTA) for
// lengths with a prime factor 7, 11 or 13.  Setup-time use only (QFS upsampling); see fft_generic.cuh.
#include "fft_generic.cuh"
#include <algorithm>

static int pick_threads(int64_t items) {
    if (items >= 2048) return 512;
    if (items >= 512) return 256;
    if (items >= 128) return 128;
    return 64;
}
template <typename K>
static int set_smem_g(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// contiguous rows, one line per CTA
__global__ void k_fft_rows_generic(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t in_pitch,
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
__global__ void k_fft_cols_generic(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t elem_stride,
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

// ------------------------------------------------------------------------------------------
// real rows
__global__ void k_r2c_rows_generic(FftPlanDev P, const cplx* __restrict__ half_tw, const double* __restrict__ in,
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

__global__ void k_c2r_rows_generic(FftPlanDev P, const cplx* __restrict__ half_tw, const cplx* __restrict__ in,
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

// out = FFT(m * IFFT(in)) / n, fused in shared memory
__global__ void k_fft_mul_fft_rows_generic(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out,
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


// ------------------------------------------------------------------------------------------
int fft_rows_generic(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
                     double scale, cudaStream_t st) {
    const int n = p->d.n;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem_g(k_fft_rows_generic, smem);
    if (rc) return rc;
    k_fft_rows_generic<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, in, out, in_pitch, out_pitch, sign, scale);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_cols_generic(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                     int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
                     cudaStream_t st) {
    const int n = p->d.n;
    int C = IPDE_FFT_GENERIC_MAX / n;
    if (C > 16) C = 16;
    if (C > n_inner) C = (int)n_inner;
    if (C < 1) C = 1;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n * C;
    int rc = set_smem_g(k_fft_cols_generic, smem);
    if (rc) return rc;
    const int64_t tiles = cdiv(n_inner, C);
    k_fft_cols_generic<<<(unsigned)(tiles * n_outer), pick_threads((int64_t)n * C / 4), smem, st>>>(
        p->d, in, out, elem_stride, n_inner, outer_stride, C, sign, scale, tw_mode, tw_big, n_big);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_r2c_rows_generic(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                         cudaStream_t st) {
    const int n = p->d.n;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem_g(k_r2c_rows_generic, smem);
    if (rc) return rc;
    k_r2c_rows_generic<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, p->half_tw_dev, in, out, in_pitch, out_pitch);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_c2r_rows_generic(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                         double scale, cudaStream_t st) {
    const int n = p->d.n;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem_g(k_c2r_rows_generic, smem);
    if (rc) return rc;
    k_c2r_rows_generic<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, p->half_tw_dev, in, out, in_pitch, out_pitch, scale);
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_mul_fft_rows_generic(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st) {
    const int n = p->d.n;
    const size_t smem = sizeof(cplx) * 2 * (size_t)n;
    int rc = set_smem_g(k_fft_mul_fft_rows_generic, smem);
    if (rc) return rc;
    k_fft_mul_fft_rows_generic<<<(unsigned)batch, pick_threads(n / 4), smem, st>>>(p->d, in, out, m);
    KERNEL_CHECK();
    return IPDE_OK;
}
