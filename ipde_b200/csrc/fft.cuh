// Hand-written FP64 Stockham FFT for sm_100a: plans + shared-memory device routines.
// No cuFFT anywhere.  See fft.cu for the kernels and DESIGN.md for the pass structure.
#pragma once
#include "common.cuh"

#define IPDE_FFT_MAX_FACTORS 24
#define IPDE_FFT_MAX_RADIX 32
// ping-pong shared-memory capacity in complex elements (2 buffers x 16 B x n <= 227 KB)
#define IPDE_FFT_SMEM_MAX 7168

struct FftPlanDev {
    int n;
    int nfac;
    int fac[IPDE_FFT_MAX_FACTORS];
    const cplx* tw;   // tw[j] = exp(-2 pi i j / n), j < n
};

struct FftPlan {
    FftPlanDev d;
    cplx* tw_dev = nullptr;
    cplx* half_tw_dev = nullptr;  // exp(-2 pi i k / (2n)), k <= n : for real transforms of length 2n
    cplx* work = nullptr;         // scratch for the global-memory path
    size_t work_elems = 0;
};

int fft_plan_create(int n, FftPlan** out);
void fft_plan_destroy(FftPlan* p);
bool fft_size_supported(int n);
int fft_next_size(int n);   // smallest even 2^a 3^b 5^c >= n

// Batched complex transforms of contiguous lines (elem stride 1), in -> out (may alias).
// sign = -1 forward, +1 inverse (unnormalised); output multiplied by `scale`.
int fft_rows(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
             int sign, double scale, cudaStream_t st);

// Batched transforms of strided lines.  Line (o, c): base = o*outer_stride + c, elements at
// base + e*elem_stride; c in [0, n_inner) are adjacent in memory.  Optional four-step twiddle:
// tw_mode 1 multiplies output element e of outer line o by exp(-/+ 2 pi i (o*e) / n_big) (sign
// follows `sign`), tw_mode 2 multiplies the *input* likewise.
int fft_cols(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
             int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
             cudaStream_t st);

// Real <-> half-complex rows (length nr = 2*p->d.n reals <-> nr/2+1 complex), plan is for nr/2.
int fft_r2c_rows(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 cudaStream_t st);
int fft_c2r_rows(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 double scale, cudaStream_t st);

// out = FFT( m * IFFT(in) ) / n  row-wise, m real (rows x n).  The annular variable-coefficient multiply
// (reference ipde/utilities.py:123-124 `ffourier_multiply`).
int fft_mul_fft_rows(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// device side
template <int R>
__device__ __forceinline__ void dft_small(cplx* v, int sign, const FftPlanDev& P);

template <>
__device__ __forceinline__ void dft_small<2>(cplx* v, int sign, const FftPlanDev& P) {
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <>
__device__ __forceinline__ void dft_small<4>(cplx* v, int sign, const FftPlanDev& P) {
    cplx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    cplx t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
    t3 = (sign < 0) ? cmulni(t3) : cmuli(t3);
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}

template <>
__device__ __forceinline__ void dft_small<8>(cplx* v, int sign, const FftPlanDev& P) {
    cplx e[4] = {v[0], v[2], v[4], v[6]};
    cplx o[4] = {v[1], v[3], v[5], v[7]};
    dft_small<4>(e, sign, P);
    dft_small<4>(o, sign, P);
    const double h = 0.70710678118654752440;
    // w8^q * o[q], w8 = exp(sign * i pi/4)
    cplx o1, o2, o3;
    if (sign < 0) {
        o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        o2 = cmulni(o[2]);
        o3 = make_double2(h * (-o[3].x + o[3].y), h * (-o[3].y - o[3].x));
    } else {
        o1 = make_double2(h * (o[1].x - o[1].y), h * (o[1].y + o[1].x));
        o2 = cmuli(o[2]);
        o3 = make_double2(h * (-o[3].x - o[3].y), h * (-o[3].y + o[3].x));
    }
    v[0] = cadd(e[0], o[0]);  v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);    v[7] = csub(e[3], o3);
}

// generic prime radix: O(R^2) with root-of-unity lookups from the plan's table
template <int R>
__device__ __forceinline__ void dft_generic(cplx* v, int sign, const FftPlanDev& P) {
    cplx w[R];
    const int step = P.n / R;
#pragma unroll
    for (int m = 0; m < R; ++m) {
        cplx t = P.tw[m * step];
        w[m] = (sign < 0) ? t : cconj(t);
    }
    cplx o[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        cplx acc = v[0];
        int idx = 0;
#pragma unroll
        for (int t = 1; t < R; ++t) {
            idx += q;
            if (idx >= R) idx -= R;
            acc = cadd(acc, cmul(v[t], w[idx]));
        }
        o[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = o[q];
}
template <> __device__ __forceinline__ void dft_small<3>(cplx* v, int s, const FftPlanDev& P) { dft_generic<3>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<5>(cplx* v, int s, const FftPlanDev& P) { dft_generic<5>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<7>(cplx* v, int s, const FftPlanDev& P) { dft_generic<7>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<11>(cplx* v, int s, const FftPlanDev& P) { dft_generic<11>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<13>(cplx* v, int s, const FftPlanDev& P) { dft_generic<13>(v, s, P); }

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
