// Generic (any supported radix, including 7/11/13) shared-memory Stockham FFT with two ping-pong buffers.
// Only used for lengths with a prime factor > 5 (e.g. the 11x upsampled boundary of the QFS setup); every
// per-solve transform goes through the tile engine of fft.cuh.
#pragma once
#include "fft.cuh"

#define IPDE_FFT_GENERIC_MAX 7168   // 2 buffers x 16 B x n <= 227 KB

// One Stockham pass of radix R over C interleaved lines held in src (element e of line c at
// src[e*C + c]); result in dst.  All threads of the CTA participate; caller syncs.
template <int R>
__device__ __forceinline__ void stockham_pass(const cplx* __restrict__ src, cplx* __restrict__ dst,
                                              const FftPlanDev& P, int Ns, int C, int sign, int tid, int nt) {
    const int n = P.n;
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const int items = m * C;
    for (int it = tid; it < items; it += nt) {
        const int c = it % C;
        const int j = it / C;
        const int k = j % Ns;
        cplx v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[(j + t * m) * C + c];
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
        for (int t = 0; t < R; ++t) dst[(j0 + t * Ns) * C + c] = v[t];
    }
}

// Full transform in shared memory; returns the buffer holding the result.  Ends with __syncthreads().
__device__ __forceinline__ cplx* stockham_smem(cplx* a, cplx* b, const FftPlanDev& P, int C, int sign, int tid,
                                               int nt) {
    int Ns = 1;
    for (int f = 0; f < P.nfac; ++f) {
        const int R = P.fac[f];
        switch (R) {
            case 2: stockham_pass<2>(a, b, P, Ns, C, sign, tid, nt); break;
            case 3: stockham_pass<3>(a, b, P, Ns, C, sign, tid, nt); break;
            case 4: stockham_pass<4>(a, b, P, Ns, C, sign, tid, nt); break;
            case 5: stockham_pass<5>(a, b, P, Ns, C, sign, tid, nt); break;
            case 7: stockham_pass<7>(a, b, P, Ns, C, sign, tid, nt); break;
            case 8: stockham_pass<8>(a, b, P, Ns, C, sign, tid, nt); break;
            case 11: stockham_pass<11>(a, b, P, Ns, C, sign, tid, nt); break;
            case 13: stockham_pass<13>(a, b, P, Ns, C, sign, tid, nt); break;
            default: break;
        }
        Ns *= R;
        __syncthreads();
        cplx* t = a; a = b; b = t;
    }
    return a;
}

int fft_rows_generic(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
                     double scale, cudaStream_t st);
int fft_cols_generic(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                     int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
                     cudaStream_t st);
int fft_r2c_rows_generic(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                         cudaStream_t st);
int fft_c2r_rows_generic(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                         double scale, cudaStream_t st);
int fft_mul_fft_rows_generic(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st);
