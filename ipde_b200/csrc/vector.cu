// Per-boundary glue of the Stokes (vector) helper: component changes and tractions on the radial grid.
//
// Device restatement of
//   EmbeddedBoundary.convert_uv_to_rt / convert_rt_to_uv   ipde/embedded_boundary.py:521-530
//   VectorHelper._get_traction_rtp / _get_traction_uvp     ipde/solvers/internals/vector.py:87-102
//   VectorHelper.__call__ (traction jumps)                 ipde/solvers/internals/vector.py:103-125
// Radial blocks are (M, n) row-major; normals/tangents are those of the boundary curve (n), broadcast over r.
#include "vector.cuh"
#include <algorithm>

#define VEC_MAX_M 40

__global__ void k_uv_to_rt(const double* __restrict__ fu, const double* __restrict__ fv, const double* __restrict__ nx,
                           const double* __restrict__ ny, const double* __restrict__ tx, const double* __restrict__ ty,
                           int M, int n, double* __restrict__ fr, double* __restrict__ ft, int inverse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = nx[i], b = ny[i], c = tx[i], d = ty[i];
    for (int r = 0; r < M; ++r) {
        const int64_t o = (int64_t)r * n + i;
        const double p = fu[o], q = fv[o];
        if (!inverse) { fr[o] = p * a + q * b; ft[o] = p * c + q * d; }        // (u, v) -> (r, t)
        else { fr[o] = p * a + q * c; ft[o] = p * b + q * d; }                  // (r, t) -> (u, v)
    }
}

int vec_uv_rt(const double* fu, const double* fv, const double* nx, const double* ny, const double* tx, const double* ty, int M,
              int n, double* out0, double* out1, int inverse, cudaStream_t st) {
    k_uv_to_rt<<<(n + 127) / 128, 128, 0, st>>>(fu, fv, nx, ny, tx, ty, M, n, out0, out1, inverse);
    KERNEL_CHECK();
    return IPDE_OK;
}

// Traction of a radial (u_r, u_t, p) field on the curve selected by `est` (a Chebyshev evaluation row), in (x, y):
//   Tr = 2 est.(D00 Ur) - est.p ;  Tt = est.(rsp * D00 (Ut/rsp)) + est.(dt Ur / rsp)     (vector.py:91-102)
// dUr = d/dt of Ur (theta derivative, computed by the caller with the FFT).
__global__ void k_radial_traction(const double* __restrict__ Ur, const double* __restrict__ Ut, const double* __restrict__ dUr,
                                  const double* __restrict__ P, const double* __restrict__ D00, const double* __restrict__ est,
                                  const double* __restrict__ rsp, const double* __restrict__ nx, const double* __restrict__ ny,
                                  const double* __restrict__ tx, const double* __restrict__ ty, int M, int n,
                                  double* __restrict__ Tx, double* __restrict__ Ty) {
    extern __shared__ double shm[];
    double* sD = shm;
    double* sE = shm + M * M;
    for (int t = threadIdx.x; t < M * M; t += blockDim.x) sD[t] = D00[t];
    for (int t = threadIdx.x; t < M; t += blockDim.x) sE[t] = est[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ur[VEC_MAX_M], w[VEC_MAX_M], s[VEC_MAX_M];
    double urt_b = 0.0, p_b = 0.0;
    for (int r = 0; r < M; ++r) {
        const int64_t o = (int64_t)r * n + i;
        s[r] = rsp[o];
        ur[r] = Ur[o];
        w[r] = Ut[o] / s[r];
        urt_b = fma(sE[r], dUr[o] / s[r], urt_b);
        p_b = fma(sE[r], P[o], p_b);
    }
    double urr_b = 0.0, utr_b = 0.0;
    for (int q = 0; q < M; ++q) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < M; ++r) { a = fma(sD[q * M + r], ur[r], a); b = fma(sD[q * M + r], w[r], b); }
        urr_b = fma(sE[q], a, urr_b);
        utr_b = fma(sE[q], s[q] * b, utr_b);
    }
    const double Tr = 2.0 * urr_b - p_b, Tt = utr_b + urt_b;
    Tx[i] = Tr * nx[i] + Tt * tx[i];
    Ty[i] = Tr * ny[i] + Tt * ty[i];
}

int vec_radial_traction(const double* Ur, const double* Ut, const double* dUr, const double* P, const double* D00,
                        const double* est, const double* rsp, const double* nx, const double* ny, const double* tx,
                        const double* ty, int M, int n, double* Tx, double* Ty, cudaStream_t st) {
    ARG_CHECK(M <= VEC_MAX_M, "M too large");
    k_radial_traction<<<(n + 127) / 128, 128, sizeof(double) * (M * M + M), st>>>(Ur, Ut, dUr, P, D00, est, rsp, nx, ny, tx, ty,
                                                                                 M, n, Tx, Ty);
    KERNEL_CHECK();
    return IPDE_OK;
}

// tau = sign * [ rtx - (txx nx + txy ny) | rty - (txy nx + tyy ny) | bu | bv ]   (vector.py:104-121)
__global__ void k_vector_jumps(const double* __restrict__ rtx, const double* __restrict__ rty, const double* __restrict__ txx,
                               const double* __restrict__ txy, const double* __restrict__ tyy, const double* __restrict__ bu,
                               const double* __restrict__ bv, const double* __restrict__ inx, const double* __restrict__ iny,
                               int n, double sign, double* __restrict__ tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double btx = txx[i] * inx[i] + txy[i] * iny[i];
    const double bty = txy[i] * inx[i] + tyy[i] * iny[i];
    tau[i] = sign * (rtx[i] - btx);
    tau[n + i] = sign * (rty[i] - bty);
    tau[2 * n + i] = sign * bu[i];
    tau[3 * n + i] = sign * bv[i];
}

int vec_jumps(const double* rtx, const double* rty, const double* txx, const double* txy, const double* tyy, const double* bu,
              const double* bv, const double* inx, const double* iny, int n, double sign, double* tau, cudaStream_t st) {
    k_vector_jumps<<<(n + 127) / 128, 128, 0, st>>>(rtx, rty, txx, txy, tyy, bu, bv, inx, iny, n, sign, tau);
    KERNEL_CHECK();
    return IPDE_OK;
}

// Pressure continuity across the interface.  A set of stokeslets determines the pressure it generates only up to what
// the velocity fixes (grad p = lap u): the grid side (grid solve + grid-side sources) and the annulus side (annular
// solve, whose constant pressure mode is free, + annulus-side sources) each carry their own additive constant, while
// the pressure of the solution is continuous across the interface.  One CTA:
//     shift = mean_j [ -(txx_j + tyy_j)/2 + pb_j - sum_i est_i pr[i, j] ],     pr += shift
// (-(txx + tyy)/2 = grid-solve pressure at interface node j, the grid velocity being divergence free; pb = pressure of
// all grid-side sources there; est = radial interpolation weights to the interface.)
__global__ void k_pressure_continuity(const double* __restrict__ txx, const double* __restrict__ tyy, const double* __restrict__ pb,
                                      const double* __restrict__ est, double* __restrict__ pr, int M, int n) {
    __shared__ double red[32];
    __shared__ double shift;
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double a = -0.5 * (txx[j] + tyy[j]) + pb[j];
        for (int i = 0; i < M; ++i) a -= est[i] * pr[(int64_t)i * n + j];
        acc += a;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) shift = v / (double)n;
    }
    __syncthreads();
    const double c = shift;
    const int64_t count = (int64_t)M * n;
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) pr[i] += c;
}
int vec_pressure_continuity(const double* txx, const double* tyy, const double* pb, const double* est, double* pr, int M, int n,
                            cudaStream_t st) {
    k_pressure_continuity<<<1, 1024, 0, st>>>(txx, tyy, pb, est, pr, M, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

// d/dt of real (M, n) rows: ifft(i k fft(f)).real with all n modes (embedded_boundary.py:473-474)
__global__ void k_real_to_cplx_rows(const double* __restrict__ in, cplx* __restrict__ out, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = make_double2(in[i], 0.0);
}
__global__ void k_mul_ik_rows(cplx* __restrict__ x, int n, int64_t total) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % n);
        const int s = (i < (n + 1) / 2) ? i : i - n;        // numpy.fft.fftfreq(n, 1/n)
        const cplx v = x[t];
        x[t] = make_double2(-(double)s * v.y, (double)s * v.x);
    }
}
__global__ void k_take_real_rows(const cplx* __restrict__ in, double* __restrict__ out, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i].x;
}

int vec_theta_derivative(FftPlan* p, const double* f, double* work_cplx, double* out, int rows, cudaStream_t st) {
    const int n = p->d.n;
    const int64_t total = (int64_t)rows * n;
    cplx* w = reinterpret_cast<cplx*>(work_cplx);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), (int64_t)ipde_num_sms() * 8));
    k_real_to_cplx_rows<<<blocks, 256, 0, st>>>(f, w, total);
    KERNEL_CHECK();
    int rc = fft_rows(p, w, w, rows, n, n, -1, 1.0, st);
    if (rc) return rc;
    k_mul_ik_rows<<<blocks, 256, 0, st>>>(w, n, total);
    KERNEL_CHECK();
    rc = fft_rows(p, w, w, rows, n, n, +1, 1.0 / n, st);
    if (rc) return rc;
    k_take_real_rows<<<blocks, 256, 0, st>>>(w, out, total);
    KERNEL_CHECK();
    return IPDE_OK;
}
