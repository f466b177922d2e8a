// Shared helpers for the ipde_b200 CUDA library (sm_100a, FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>

#define IPDE_OK 0
#define IPDE_ERR_CUDA 1
#define IPDE_ERR_ARG 2
#define IPDE_ERR_UNSUPPORTED 3

void ipde_set_error(const char* fmt, ...);
void ipde_count_launch();   // every kernel launch of this library is counted (bench.py: gpu_launches)
// Per-kernel timing (off unless ipde_profile_start was called): KERNEL_CHECK records a CUDA event after the
// launch on `prof_stream`; KERNEL_TAG names the launch and states its algorithmic bytes (bench.py: roofline).
void ipde_prof_tag(const char* name, double bytes);
void ipde_prof_mark(const char* file, int line);
#define KERNEL_TAG(name, bytes) ipde_prof_tag(name, (double)(bytes))

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ipde_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                           cudaGetErrorString(_e));                                 \
            return IPDE_ERR_CUDA;                                                   \
        }                                                                           \
    } while (0)

#define KERNEL_CHECK()                  \
    do {                                \
        ipde_count_launch();            \
        ipde_prof_mark(__FILE__, __LINE__); \
        CUDA_TRY(cudaGetLastError());   \
    } while (0)

#define ARG_CHECK(cond, msg)                                                        \
    do {                                                                            \
        if (!(cond)) {                                                              \
            ipde_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, msg);     \
            return IPDE_ERR_ARG;                                                    \
        }                                                                           \
    } while (0)

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cmuli(cplx a) { return make_double2(-a.y, a.x); }    // i*a
__host__ __device__ __forceinline__ cplx cmulni(cplx a) { return make_double2(a.y, -a.x); }   // -i*a

static inline int ipde_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
