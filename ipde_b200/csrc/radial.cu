// Radial (Chebyshev-Gauss x periodic) block -> arbitrary points (r, t) inside the annulus.
//
// Device restatement of EmbeddedBoundary.interpolate_radial_to_points (ipde/embedded_boundary.py:419-434):
// the reference even-extends the M rows to 2M, takes a 2-D FFT, applies a half-sample phase, zeroes the
// radial Nyquist row and runs a 2-D type-2 NUFFT at (theta, t), theta = arccos of the affinely mapped r.
// That sum is exactly  sum_{m<M} a_m(t) cos(m theta)  with a_m(t) the trigonometric interpolant (Nyquist
// taken as a cosine) of the discrete cosine coefficients of each column, so here:
//   1. coef = Cmat (M x M) @ fr           -- DCT over r, tiny dense product
//   2. every coefficient row is oversampled in t by our 1-D NUFFT preparation (nufft.cu)
//   3. one warp per target: ES-kernel interpolation in t for all M rows and a Chebyshev recurrence in r.
#include "nufft.cuh"
#include "radial.cuh"
#include <algorithm>

static const double TWO_PI = 6.283185307179586476925286766559;

__global__ void k_small_matmul(const double* __restrict__ C, const double* __restrict__ fr, double* __restrict__ out, int M, int n) {
    extern __shared__ double sC[];
    for (int t = threadIdx.x; t < M * M; t += blockDim.x) sC[t] = C[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int q = 0; q < M; ++q) {
        double acc = 0.0;
        for (int r = 0; r < M; ++r) acc = fma(sC[q * M + r], fr[(int64_t)r * n + i], acc);
        out[(int64_t)q * n + i] = acc;
    }
}

// fine: (M, fn) complex oversampled coefficient rows (real part used)
__global__ void k_radial_interp(const cplx* __restrict__ fine, int M, int fn, int w, double beta, const double* __restrict__ theta,
                                const double* __restrict__ t, int64_t npts, double* __restrict__ out,
                                const int64_t* __restrict__ out_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double half = 0.5 * w;
    for (int64_t p = warp; p < npts; p += nwarps) {
        const double X = t[p] * ((double)fn / TWO_PI);
        const int i0 = (int)ceil(X - half);
        const double x = cos(theta[p]);
        // lane -> kernel tap a = lane % 16 (and a + 16 if w > 16), row parity g = lane / 16
        const int g = lane >> 4;
        double acc = 0.0;
        for (int a = lane & 15; a < w; a += 16) {
            const double z = ((double)(i0 + a) - X) / half;
            const double arg = 1.0 - z * z;
            if (arg < 0.0) continue;
            const double k = exp(beta * (sqrt(arg) - 1.0));
            int ii = (i0 + a) % fn; if (ii < 0) ii += fn;
            // Chebyshev recurrence T_m(x); this lane takes rows m = g, g+2, ...
            double T0 = 1.0, T1 = x;
            double s = 0.0;
            for (int m = 0; m < M; ++m) {
                double T;
                if (m == 0) T = T0;
                else if (m == 1) T = T1;
                else { T = 2.0 * x * T1 - T0; T0 = T1; T1 = T; }
                if ((m & 1) == g) s = fma(T, fine[(int64_t)m * fn + ii].x, s);
            }
            acc = fma(k, s, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[out_pos ? out_pos[p] : p] = acc;
    }
}

int radial_to_points(Nufft1Plan* plan, const double* cmat, int M, const double* fr, double* coef_work,
                     const double* theta, const double* t, int64_t npts, double* out, const int64_t* out_pos,
                     cudaStream_t st) {
    const int n = plan->n;
    k_small_matmul<<<(n + 127) / 128, 128, sizeof(double) * M * M, st>>>(cmat, fr, coef_work, M, n);
    KERNEL_CHECK();
    int rc = nufft1_upsample_rows(plan, coef_work, M, st);
    if (rc) return rc;
    if (npts > 0) {
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(npts * 32, 256), (int64_t)ipde_num_sms() * 16));
        k_radial_interp<<<blocks, 256, 0, st>>>(plan->buf_f, M, plan->fn, plan->w, plan->beta, theta, t, npts, out, out_pos);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}
