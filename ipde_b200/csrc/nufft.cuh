#pragma once
#include "grid.cuh"

struct Nufft2Plan {
    int nx, ny, fx, fy, w;
    double beta, xh, yh;
    GridPlan* fine = nullptr;
    double* px = nullptr;       // deconvolution factors per coarse x frequency index (FFT order)
    double* py = nullptr;       // ... per coarse y frequency index 0..ny/2
    cplx* fine_spec = nullptr;
    double* fine_real = nullptr;
    uint32_t* mask = nullptr;   // bit per (fine row, block of 64 reals) read by the current targets
};

int nufft2_plan_create(int nx, int ny, double xh, double yh, int fx, int fy, int w, double beta, const double* px_host,
                       const double* py_host, Nufft2Plan** out);
void nufft2_plan_destroy(Nufft2Plan* p);
// spec: coarse half spectrum (rows possibly in four-step order cn1 x cn2); dxp/dyp: multiply by (i kx), (i ky)
int nufft2_apply(Nufft2Plan* p, const cplx* spec, int cn1, int cn2, int dxp, int dyp, const double* x, const double* y,
                 int64_t npts, double* out, cudaStream_t st);

// value, d/dx, d/dy at the targets from one fine-grid transform (kernel differentiation)
int nufft2_apply3(Nufft2Plan* p, const cplx* spec, int cn1, int cn2, const double* x, const double* y, int64_t npts,
                  double* out_v, double* out_dx, double* out_dy, cudaStream_t st);

struct Nufft1Plan {
    int n, fn, w;
    double beta;
    FftPlan* coarse = nullptr;
    FftPlan* fine = nullptr;
    double* dec = nullptr;
    cplx* buf_c = nullptr;
    cplx* buf_f = nullptr;      // (rows, fn) oversampled rows after nufft1_upsample_rows
    int cap_rows = 0;
};
int nufft1_plan_create(int n, int fn, int w, double beta, const double* p_host, Nufft1Plan** out);
void nufft1_plan_destroy(Nufft1Plan* p);
int nufft1_upsample_rows(Nufft1Plan* p, const double* rows_real, int nrows, cudaStream_t st);
