// FP64 modified Bessel functions K0, K1 for the device (CUDA's math library has neither).
// x <= 2: power series; x > 2: exp(-x)/sqrt(x) * Chebyshev series in t = 4/x - 1, coefficients fitted on the
// host against scipy.special.k0e / k1e and uploaded once (ipde_set_k0_table).
#pragma once
#include "common.cuh"
#define K0_MAX_CHEB 40
extern __constant__ double c_k0_cheb[K0_MAX_CHEB];
extern __constant__ double c_k1_cheb[K0_MAX_CHEB];
extern __constant__ int c_k0_ncheb;
extern __constant__ double c_k0_a[16];   // 1/(m!)^2
extern __constant__ double c_k0_b[16];   // H_m/(m!)^2
extern __constant__ double c_k1_a[16];   // 1/(m!(m+1)!)
extern __constant__ double c_k1_b[16];   // (H_m + H_{m+1})/(2 m!(m+1)!)  (with the -gamma folded in at use)
bool bessel_tables_ready();

__device__ __forceinline__ double cheb_tail(double x, const double* c, int nc) {
    const double t = 4.0 / x - 1.0, t2 = 2.0 * t;
    double b0 = 0.0, b1 = 0.0;
    for (int m = nc - 1; m >= 1; --m) {
        const double tmp = fma(t2, b0, c[m]) - b1;
        b1 = b0;
        b0 = tmp;
    }
    return fma(t, b0, c[0]) - b1;
}

// K0(x) for x > 0, given x and x^2
__device__ __forceinline__ double bessel_k0(double x, double x2) {
    if (x <= 2.0) {
        const double y = 0.25 * x2;
        double i0 = c_k0_a[15], s = c_k0_b[15];
#pragma unroll
        for (int m = 14; m >= 0; --m) {
            i0 = fma(i0, y, c_k0_a[m]);
            s = fma(s, y, c_k0_b[m]);
        }
        const double EULER = 0.57721566490153286061;
        return -(0.5 * log(y) + EULER) * i0 + s;
    }
    return exp(-x) * rsqrt(x) * cheb_tail(x, c_k0_cheb, c_k0_ncheb);
}

// K1(x) for x > 0:  K1 = 1/x + (x/2) [ (ln(x/2) + gamma) I1(x)/(x/2) - sum_m (x^2/4)^m (H_m + H_{m+1})/(2 m!(m+1)!) ]
__device__ __forceinline__ double bessel_k1(double x, double x2) {
    if (x <= 2.0) {
        const double y = 0.25 * x2;
        double i1 = c_k1_a[15], s = c_k1_b[15];
#pragma unroll
        for (int m = 14; m >= 0; --m) {
            i1 = fma(i1, y, c_k1_a[m]);
            s = fma(s, y, c_k1_b[m]);
        }
        const double EULER = 0.57721566490153286061;
        return 1.0 / x + 0.5 * x * ((0.5 * log(y) + EULER) * i1 - s);
    }
    return exp(-x) * rsqrt(x) * cheb_tail(x, c_k1_cheb, c_k0_ncheb);
}
