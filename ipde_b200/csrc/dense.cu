// Dense and sparse linear-algebra kernels used per solve:
//   * row-major GEMV (QFS boundary->check operators; reference: qfs package, applied at
//     ipde/solvers/internals/scalar.py:87-88)
//   * LU substitution with the factors of the check<-source matrix (two-step QFS apply and u2s,
//     internals/scalar.py:106), blocked, one cooperative kernel per triangular factor
//   * CSR SpMV, warp per row (reference ipde/sparse_matvec.py:17-113, MKL mkl_cspblas_{z,d}csrgemv)
//   * small fused vector helpers for the per-boundary glue of ScalarHelper.__call__ (internals/scalar.py:73-85)
#include "dense.cuh"
#include <cooperative_groups.h>
#include <algorithm>
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------
__global__ void k_gemv(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y, int64_t m,
                       int64_t n, int64_t lda) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec2 = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t r = warp; r < m; r += nwarps) {
        const double* a = A + r * lda;
        double acc0 = 0.0, acc1 = 0.0;
        int64_t c = 0;
        if (vec2) {
            // 16-byte loads: 4 independent double2 loads of the row per lane per iteration (2 KB per warp)
            const double2* a2 = reinterpret_cast<const double2*>(a);
            const double2* x2 = reinterpret_cast<const double2*>(x);
            const int64_t n2 = n >> 1;
            int64_t d = lane;
            for (; d + 96 < n2; d += 128) {
                const double2 p0 = a2[d], p1 = a2[d + 32], p2 = a2[d + 64], p3 = a2[d + 96];
                const double2 q0 = x2[d], q1 = x2[d + 32], q2 = x2[d + 64], q3 = x2[d + 96];
                acc0 = fma(p0.x, q0.x, acc0); acc1 = fma(p0.y, q0.y, acc1);
                acc0 = fma(p1.x, q1.x, acc0); acc1 = fma(p1.y, q1.y, acc1);
                acc0 = fma(p2.x, q2.x, acc0); acc1 = fma(p2.y, q2.y, acc1);
                acc0 = fma(p3.x, q3.x, acc0); acc1 = fma(p3.y, q3.y, acc1);
            }
            for (; d < n2; d += 32) {
                const double2 p0 = a2[d], q0 = x2[d];
                acc0 = fma(p0.x, q0.x, acc0); acc1 = fma(p0.y, q0.y, acc1);
            }
            c = n2 << 1;
        }
        for (c += lane; c < n; c += 32) acc0 = fma(a[c], x[c], acc0);
        double acc = acc0 + acc1;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[r] = acc;
    }
}

int dense_gemv(const double* A, const double* x, double* y, int64_t m, int64_t n, cudaStream_t st) {
    if (m <= 0) return IPDE_OK;
    const int threads = 256;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(m * 32, threads), (int64_t)ipde_num_sms() * 8));
    KERNEL_TAG("k_gemv", 8.0 * (double)m * (double)n);
    k_gemv<<<blocks, threads, 0, st>>>(A, x, y, m, n, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Blocked triangular solves.  LU is row-major (M x M): unit-lower L below the diagonal, U on and above.
#define TRI_B 64
#define TRI_THREADS 256

// x <- L^{-1} x   (forward), every CTA solves the diagonal block redundantly, then updates its share of rows
__global__ void __launch_bounds__(TRI_THREADS) k_trsv_lower_unit(const double* __restrict__ LU, double* __restrict__ x, int M) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double D[TRI_B][TRI_B + 1];
    __shared__ double y[TRI_B];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * warps_per_block + warp_in_block;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    for (int j0 = 0; j0 < M; j0 += TRI_B) {
        const int bs = min(TRI_B, M - j0);
        for (int t = threadIdx.x; t < bs * bs; t += blockDim.x) {
            const int r = t / bs, c = t % bs;
            D[r][c] = LU[(int64_t)(j0 + r) * M + j0 + c];
        }
        if (threadIdx.x < bs) y[threadIdx.x] = x[j0 + threadIdx.x];
        __syncthreads();
        for (int i = 0; i < bs - 1; ++i) {
            const double yi = y[i];
            const int r = i + 1 + threadIdx.x;
            if (r < bs) y[r] -= D[r][i] * yi;
            __syncthreads();
        }
        // update the remaining rows
        const int r0 = j0 + bs;
        for (int64_t r = r0 + gwarp; r < M; r += nwarps) {
            const double* a = LU + r * M + j0;
            double acc = 0.0;
            for (int c = lane; c < bs; c += 32) acc = fma(a[c], y[c], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) x[r] -= acc;
        }
        grid.sync();
        // every CTA has read the unsolved block by now: publish the solved one
        if (blockIdx.x == 0 && threadIdx.x < bs) x[j0 + threadIdx.x] = y[threadIdx.x];
    }
}

// x <- U^{-1} x   (backward)
__global__ void __launch_bounds__(TRI_THREADS) k_trsv_upper(const double* __restrict__ LU, double* __restrict__ x, int M) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double D[TRI_B][TRI_B + 1];
    __shared__ double y[TRI_B];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * warps_per_block + warp_in_block;
    const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
    const int nblk = (M + TRI_B - 1) / TRI_B;
    for (int jb = nblk - 1; jb >= 0; --jb) {
        const int j0 = jb * TRI_B;
        const int bs = min(TRI_B, M - j0);
        for (int t = threadIdx.x; t < bs * bs; t += blockDim.x) {
            const int r = t / bs, c = t % bs;
            D[r][c] = LU[(int64_t)(j0 + r) * M + j0 + c];
        }
        if (threadIdx.x < bs) y[threadIdx.x] = x[j0 + threadIdx.x];
        __syncthreads();
        for (int i = bs - 1; i >= 0; --i) {
            if (threadIdx.x == 0) y[i] /= D[i][i];
            __syncthreads();
            const double yi = y[i];
            if ((int)threadIdx.x < i) y[threadIdx.x] -= D[threadIdx.x][i] * yi;
            __syncthreads();
        }
        for (int64_t r = gwarp; r < j0; r += nwarps) {
            const double* a = LU + r * M + j0;
            double acc = 0.0;
            for (int c = lane; c < bs; c += 32) acc = fma(a[c], y[c], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) x[r] -= acc;
        }
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x < bs) x[j0 + threadIdx.x] = y[threadIdx.x];
    }
}

__global__ void k_permute(const double* __restrict__ b, const int* __restrict__ perm, double* __restrict__ x, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) x[i] = b[perm[i]];
}

static int coop_grid(const void* kernel) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRI_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    return per_sm * ipde_num_sms();
}

int dense_lu_solve(const double* LU, const int* perm, const double* b, double* x, int M, cudaStream_t st) {
    if (M <= 0) return IPDE_OK;
    k_permute<<<(M + 255) / 256, 256, 0, st>>>(b, perm, x, M);
    KERNEL_CHECK();
    void* args[] = {(void*)&LU, (void*)&x, (void*)&M};
    int g1 = std::min(coop_grid((const void*)k_trsv_lower_unit), std::max(1, (M + TRI_B - 1) / TRI_B * 2));
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_trsv_lower_unit, dim3(g1), dim3(TRI_THREADS), args, 0, st));
    ipde_count_launch();
    int g2 = std::min(coop_grid((const void*)k_trsv_upper), std::max(1, (M + TRI_B - 1) / TRI_B * 2));
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_trsv_upper, dim3(g2), dim3(TRI_THREADS), args, 0, st));
    ipde_count_launch();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// CSR SpMV, one warp per row
template <typename T>
__device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ cplx zero_of<cplx>() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double madd(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ cplx madd(cplx a, cplx b, cplx c) { return cadd(c, cmul(a, b)); }
__device__ __forceinline__ double wsum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx wsum(cplx v) {
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <typename T>
__global__ void k_csr_spmv(const int* __restrict__ indptr, const int* __restrict__ indices, const T* __restrict__ data,
                           const T* __restrict__ x, T* __restrict__ y, int64_t nrows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrows; r += nwarps) {
        const int a = indptr[r], b = indptr[r + 1];
        T acc = zero_of<T>();
        for (int e = a + lane; e < b; e += 32) acc = madd(data[e], x[indices[e]], acc);
        acc = wsum(acc);
        if (lane == 0) y[r] = acc;
    }
}

int csr_spmv_d(const int* indptr, const int* indices, const double* data, const double* x, double* y, int64_t nrows,
               cudaStream_t st) {
    if (nrows <= 0) return IPDE_OK;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(nrows * 32, 256), (int64_t)ipde_num_sms() * 8));
    k_csr_spmv<double><<<blocks, 256, 0, st>>>(indptr, indices, data, x, y, nrows);
    KERNEL_CHECK();
    return IPDE_OK;
}
int csr_spmv_z(const int* indptr, const int* indices, const cplx* data, const cplx* x, cplx* y, int64_t nrows,
               cudaStream_t st) {
    if (nrows <= 0) return IPDE_OK;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(nrows * 32, 256), (int64_t)ipde_num_sms() * 8));
    k_csr_spmv<cplx><<<blocks, 256, 0, st>>>(indptr, indices, data, x, y, nrows);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// small vector helpers
__global__ void k_axpby(double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* __restrict__ out,
                        int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
__global__ void k_mul(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = x[i] * y[i];
}
static inline int vblocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, 256), 2048)); }

int vec_axpby(double a, const double* x, double b, const double* y, double* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_axpby<<<vblocks(n), 256, 0, st>>>(a, x, b, y, out, n);
    KERNEL_CHECK();
    return IPDE_OK;
}
int vec_mul(const double* x, const double* y, double* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_mul<<<vblocks(n), 256, 0, st>>>(x, y, out, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

// out[i] = sum_r row[r] * fr[r*n + i]     (1 x M Chebyshev row applied to an (M, n) radial block)
__global__ void k_rowvec(const double* __restrict__ row, const double* __restrict__ fr, double* __restrict__ out, int M, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < M; ++r) acc = fma(row[r], fr[(int64_t)r * n + i], acc);
        out[i] = acc;
    }
}
int rowvec_apply(const double* row, const double* fr, double* out, int M, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_rowvec<<<vblocks(n), 256, 0, st>>>(row, fr, out, M, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ScalarHelper.__call__ glue (internals/scalar.py:73-85):
//   tau[0:n] = sign*(in_est . ur - (bx nx + by ny)) ; tau[n:2n] = sign*bv
__global__ void k_scalar_jumps(const double* __restrict__ bv, const double* __restrict__ bx, const double* __restrict__ by,
                               const double* __restrict__ nx, const double* __restrict__ ny, const double* __restrict__ ur,
                               const double* __restrict__ in_est, int M, int64_t n, double sign, double* __restrict__ tau) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double urn = 0.0;
        for (int r = 0; r < M; ++r) urn = fma(in_est[r], ur[(int64_t)r * n + i], urn);
        const double ucn = bx[i] * nx[i] + by[i] * ny[i];
        tau[i] = sign * (urn - ucn);
        tau[n + i] = sign * bv[i];
    }
}
int scalar_jumps(const double* bv, const double* bx, const double* by, const double* nx, const double* ny, const double* ur,
                 const double* in_est, int M, int64_t n, double sign, double* tau, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_scalar_jumps<<<vblocks(n), 256, 0, st>>>(bv, bx, by, nx, ny, ur, in_est, M, n, sign, tau);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Structured QFS apply (ipde_b200/qfs_structured.py; replaces the dense factors of the reference's qfs objects used
// at ipde/solvers/internals/scalar.py:87-88,106 and internals/vector.py:124-125,150):
//   out_hat_o[k] = sum_i ( sum_{|p|<=BH} band[o,i,p,k] in_i[k-p]  +  [|k|<=KL] sum_{|k'|<=KL} low[o,i,k,k'] in_i[k'] )
// in_i[k'] = coefficient of exp(i k' t) of the i-th density (N-point FFT / N, Nyquist mode split evenly);
// output modes |k| <= Ks are written into a length-M spectrum in FFT order (zero elsewhere); two real outputs share
// one complex spectrum (out_{2q} + i out_{2q+1}), so a single inverse transform returns both.
__device__ __forceinline__ cplx banded_in_mode(const cplx* __restrict__ xh, int64_t N, int64_t h, bool even, int64_t kp,
                                               double invN) {
    if (kp < -h || kp > h) return make_double2(0.0, 0.0);
    const cplx v = xh[kp >= 0 ? kp : kp + N];
    const double s = (even && (kp == h || kp == -h)) ? 0.5 * invN : invN;
    return make_double2(v.x * s, v.y * s);
}

// The operators map real inputs to real outputs (Q(-k,-k') = conj Q(k,k'), enforced at setup), so only the output
// modes k >= 0 are computed -- band[o,i,p,k] and lowT[o,i,k',k] are stored for k in [0, Ks] / [0, KL] only -- and
// mode -k of every real output is the conjugate.
template <int PS>
__global__ void __launch_bounds__(32 * PS)
k_banded_apply(const cplx* __restrict__ xh, int n_in, int64_t N, const cplx* __restrict__ band,
               const cplx* __restrict__ lowT, int n_out, int BH, int KL, int64_t Ks, cplx* __restrict__ yh, int64_t M) {
    __shared__ double red[PS][32][4];
    const int lane = threadIdx.x, ps = threadIdx.y;
    const int64_t k = blockIdx.x * 32LL + lane;             // output mode, 0 .. M/2
    const int o0 = 2 * blockIdx.y, o1 = o0 + 1;
    const bool two = o1 < n_out;
    const bool live = k <= Ks;
    const int64_t nk = Ks + 1;
    const int64_t h = N / 2;
    const bool even = (N % 2) == 0;
    const int64_t hh = even ? h : (N - 1) / 2;
    const double invN = 1.0 / (double)N;
    const int nb = 2 * BH + 1, nl = 2 * KL + 1, nlk = KL + 1;
    cplx a0 = make_double2(0.0, 0.0), a1 = a0;
    if (live) {
        for (int i = 0; i < n_in; ++i) {
            const cplx* x = xh + (int64_t)i * N;
            const cplx* b0 = band + ((int64_t)(o0 * n_in + i) * nb) * nk + k;
            const cplx* b1 = band + ((int64_t)((two ? o1 : o0) * n_in + i) * nb) * nk + k;
            for (int p = -BH + ps; p <= BH; p += PS) {
                const cplx v = banded_in_mode(x, N, hh, even, k - p, invN);
                const cplx c0 = b0[(int64_t)(p + BH) * nk];
                a0.x = fma(c0.x, v.x, fma(-c0.y, v.y, a0.x));
                a0.y = fma(c0.x, v.y, fma(c0.y, v.x, a0.y));
                if (two) {
                    const cplx c1 = b1[(int64_t)(p + BH) * nk];
                    a1.x = fma(c1.x, v.x, fma(-c1.y, v.y, a1.x));
                    a1.y = fma(c1.x, v.y, fma(c1.y, v.x, a1.y));
                }
            }
            if (k <= KL) {
                const cplx* l0 = lowT + ((int64_t)(o0 * n_in + i) * nl) * nlk + k;
                const cplx* l1 = lowT + ((int64_t)((two ? o1 : o0) * n_in + i) * nl) * nlk + k;
                for (int kp = -KL + ps; kp <= KL; kp += PS) {
                    const cplx v = banded_in_mode(x, N, hh, even, kp, invN);
                    const cplx c0 = l0[(int64_t)(kp + KL) * nlk];
                    a0.x = fma(c0.x, v.x, fma(-c0.y, v.y, a0.x));
                    a0.y = fma(c0.x, v.y, fma(c0.y, v.x, a0.y));
                    if (two) {
                        const cplx c1 = l1[(int64_t)(kp + KL) * nlk];
                        a1.x = fma(c1.x, v.x, fma(-c1.y, v.y, a1.x));
                        a1.y = fma(c1.x, v.y, fma(c1.y, v.x, a1.y));
                    }
                }
            }
        }
    }
    red[ps][lane][0] = a0.x; red[ps][lane][1] = a0.y; red[ps][lane][2] = a1.x; red[ps][lane][3] = a1.y;
    __syncthreads();
    if (ps == 0 && k <= M / 2) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int q = 0; q < PS; ++q) { s0 += red[q][lane][0]; s1 += red[q][lane][1]; s2 += red[q][lane][2]; s3 += red[q][lane][3]; }
        cplx* y = yh + (int64_t)blockIdx.y * M;
        // mode k: a0 + i a1;  mode -k: conj(a0) + i conj(a1)
        y[k] = make_double2(s0 - s3, s1 + s2);
        if (k > 0 && k < M - k) y[M - k] = make_double2(s0 + s3, s2 - s1);
    }
}

int banded_apply_z(const cplx* xh, int n_in, int64_t N, const cplx* band, const cplx* lowT, int n_out, int BH, int KL,
                   int64_t Ks, cplx* yh, int64_t M, cudaStream_t st) {
    ARG_CHECK(n_in >= 1 && n_out >= 1 && N >= 2 && M >= 2 * Ks + 2 && BH >= 0 && KL >= 0 && KL <= Ks, "banded_apply_z sizes");
    constexpr int PS = 8;
    dim3 grid((unsigned)cdiv(M / 2 + 1, 32), (unsigned)((n_out + 1) / 2));
    KERNEL_TAG("k_banded_apply", 16.0 * ((double)n_out * n_in * ((2.0 * BH + 1) * (Ks + 1.0) + (2.0 * KL + 1) * (KL + 1.0))));
    k_banded_apply<PS><<<grid, dim3(32, PS), 0, st>>>(xh, n_in, N, band, lowT, n_out, BH, KL, Ks, yh, M);
    KERNEL_CHECK();
    return IPDE_OK;
}

__global__ void k_real_to_complex(const double* __restrict__ x, cplx* __restrict__ z, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = make_double2(x[i], 0.0);
}
int real_to_complex(const double* x, cplx* z, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_real_to_complex<<<vblocks(n), 256, 0, st>>>(x, z, n);
    KERNEL_CHECK();
    return IPDE_OK;
}
// re[i] = Re z[i] (* wr[i]), im[i] = Im z[i] (* wi[i]); im / weights may be null
__global__ void k_complex_parts(const cplx* __restrict__ z, double* __restrict__ re, double* __restrict__ im, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const cplx v = z[i];
        re[i] = v.x;
        if (im) im[i] = v.y;
    }
}
int complex_parts(const cplx* z, double* re, double* im, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_complex_parts<<<vblocks(n), 256, 0, st>>>(z, re, im, n);
    KERNEL_CHECK();
    return IPDE_OK;
}
// re[j] (+)= Re z[j*stride], im[j] (+)= Im z[j*stride]  (curve operators: the field is transformed back on the
// source grid of M = stride*N points and every stride-th sample is a target node)
__global__ void k_complex_parts_strided(const cplx* __restrict__ z, double* __restrict__ re, double* __restrict__ im,
                                        int64_t n, int64_t stride, int accumulate) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const cplx v = z[i * stride];
        if (accumulate) { re[i] += v.x; if (im) im[i] += v.y; }
        else { re[i] = v.x; if (im) im[i] = v.y; }
    }
}
int complex_parts_strided(const cplx* z, double* re, double* im, int64_t n, int64_t stride, int accumulate, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_complex_parts_strided<<<vblocks(n), 256, 0, st>>>(z, re, im, n, stride, accumulate);
    KERNEL_CHECK();
    return IPDE_OK;
}
