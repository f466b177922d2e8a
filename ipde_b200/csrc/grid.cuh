#pragma once
#include "fft.cuh"
#include "fft_short.cuh"

#define IPDE_KIND_POISSON 0
#define IPDE_KIND_MODHELM 1
#define IPDE_KIND_STOKES 2
#define IPDE_KIND_LAPLACIAN 3

// Half-spectrum layout of a real (nx, ny) grid function: complex (nx, ny/2+1), row pitch ny/2+1.
// When the column transform is split (n2 > 1) the rows of a spectrum are in four-step order:
// frequency index k1 + n1*k2 is stored at row n2*k1 + k2.
struct GridPlan {
    int nx, ny;
    double xh, yh;
    int64_t pitch;
    int n1, n2;
    FftPlan* row = nullptr;    // length ny/2 (packed real rows)
    FftPlan* colA = nullptr;   // length n1
    FftPlan* colB = nullptr;   // length n2 (four-step only)
    cplx* tw_big = nullptr;    // exp(-2 pi i j / nx)
    double* reduce_buf = nullptr;
};

int grid_plan_create(int nx, int ny, double xh, double yh, GridPlan** out);
void grid_plan_destroy(GridPlan* g);
int grid_cols_forward(GridPlan* g, const cplx* in, cplx* out, cudaStream_t st);
int grid_cols_inverse(GridPlan* g, const cplx* in, cplx* out, cudaStream_t st);
int grid_cols_inverse_n(GridPlan* g, const cplx* in, cplx* out, int64_t ncols, cudaStream_t st);
int grid_fft2_forward(GridPlan* g, const double* real_in, cplx* spec, cudaStream_t st);
int grid_fft2_inverse(GridPlan* g, const cplx* spec_in, cplx* spec_work, double* real_out, cudaStream_t st);
int grid_scatter_cutoff(GridPlan* g, const double* f_compact, const int64_t* phys_lin, const double* step_compact,
                        int64_t n_phys, double* fc, cudaStream_t st);
int grid_demean(GridPlan* g, double* f, const double* bumpy, cudaStream_t st);
int grid_scalar_multiply(GridPlan* g, cplx* spec, int kind, double k, cudaStream_t st);
int grid_fd4(GridPlan* g, const double* f, double* fx, double* fy, cudaStream_t st);
int grid_interp_cubic(GridPlan* g, const double* f, double x0, double y0, const double* px, const double* py,
                      int64_t npts, double* out, cudaStream_t st);
int grid_gather(const double* uc, const int64_t* lin, int64_t n, double* out, cudaStream_t st);
int grid_index_add(double* out, const int64_t* pos, const double* v, int64_t n, cudaStream_t st);
int grid_index_set(double* out, const int64_t* pos, const double* v, int64_t n, cudaStream_t st);
int grid_stokes_multiply(GridPlan* g, cplx* uh, cplx* vh, cplx* ph, cudaStream_t st);

// real symbol applied between the forward and inverse column transforms (grid.cu: k_cols_mult)
struct ColMult {
    int kind;             // IPDE_KIND_POISSON, IPDE_KIND_MODHELM (analytic) or 3 (table)
    double k2c;           // k^2 of the modified Helmholtz operator
    int nx, ny;
    double xh, yh;
    int fmul;             // frequency index along x of element e of outer line o: o + fmul*e
    const double* table;  // kind 3: one real per spectrum entry, same layout as the spectrum
    int col_offset;       // global index of column 0 (slab-decomposed grids hold a block of columns)
    int deriv;            // kind 3 only: 0 plain, 1 times i kx, 2 times i ky (Nyquist row/column -> 0)
};
// win (optional): {in_lo, in_hi, out_lo, out_hi} row windows of zero-padded data, see RowWindow in fft.cuh
int grid_cols_symbol(GridPlan* g, cplx* spec_in, cplx* spec_out, cplx* spec_copy, int copy_mode, const ColMult& M,
                     cudaStream_t st, const int* win = nullptr, const SlabScatter* sc = nullptr);
// the same on the column slab of one rank, the last pass storing straight into the row slabs of all ranks
int grid_cols_symbol_slab_scatter(GridPlan* g, cplx* spec_in, cplx* spec_work, int kind, double k, int ny_global, double yh,
                                  int col_offset, const SlabScatter& sc, cudaStream_t st);
// win (optional): {out_lo, out_hi}
int grid_cols_symbol_inverse(GridPlan* g, const cplx* spec2d, cplx* spec_out, const ColMult& M, cudaStream_t st,
                             const int* win = nullptr);
int grid_solve_scalar_fused(GridPlan* g, int kind, double k, const double* fc, cplx* uch, cplx* spec_work, double* uc,
                            cudaStream_t st);
int grid_stokes_stress(GridPlan* g, const cplx* uh, const cplx* vh, const cplx* ph, cplx* txx, cplx* txy, cplx* tyy,
                       cudaStream_t st);

// ---- slab decomposition (ipde_b200/distributed.py): a rank holds `rows` rows of a half spectrum of pitch P; the
// all-to-all transpose sends column block q (Pc columns, zero padded past P) to rank q.
int grid_slab_pack(const cplx* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, cplx* out,
                   cudaStream_t st);
int grid_slab_unpack(const cplx* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, cplx* out,
                     cudaStream_t st);
// column pass of a column slab: plan g has nx = global nx and pitch == Pc; ky is taken from the global grid
int grid_cols_symbol_slab(GridPlan* g, cplx* spec_in, cplx* spec_out, cplx* spec_copy, int copy_mode, int kind, double k,
                          int ny_global, double yh, int col_offset, cudaStream_t st);
