/*
 * libipde_b200 -- C ABI of the B200-native (sm_100a, FP64) kernels behind ipde's
 * inhomogeneous-solve hot path.
 *
 * The reference (dbstein/ipde) is pure Python; it has no FFI for this path.  The seams at
 * which native code sits in the reference -- and which these entry points replace -- are
 * cited per function as `reference file:line`.  Callers: the Python solver classes in
 * ipde_b200/solvers (ctypes, see ipde_b200/_lib.py); INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on error; ipde_last_error() gives the text;
 *   - all array pointers are DEVICE pointers owned by the caller unless marked `host`;
 *   - complex data is interleaved (re, im) doubles; sizes are int64_t; `stream` is a cudaStream_t;
 *   - plans are opaque handles created/destroyed by the library; a plan is not thread-safe,
 *     different plans may be used from different threads;
 *   - grids are (nx, ny) row-major with x the slow index (numpy meshgrid indexing='ij').
 */
#ifndef IPDE_B200_H
#define IPDE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPDE_KIND_POISSON 0
#define IPDE_KIND_MODHELM 1
#define IPDE_KIND_STOKES 2
#define IPDE_KIND_LAPLACIAN 3   /* ipde_grid_scalar_multiply only: the symbol -(kx^2 + ky^2) itself (ebdyc.laplacian) */

const char* ipde_last_error(void);
int ipde_version(void);
int ipde_device_sm_count(void);
int64_t ipde_launch_count(void);   /* kernels launched by this library since load */
/* per-kernel CUDA-event timing of everything launched on `stream` between start and stop; stop writes lines
 * "kernel_name launches total_ms algorithmic_bytes" (sorted by time) into the host buffer `out` */
int ipde_profile_start(void* stream);
int ipde_profile_stop(char* out, int64_t cap);

/* ---- 1-D FFT (batched, contiguous lines).  Replaces numpy.fft.fft / ifft, ipde/utilities.py:5-12. */
int ipde_fft_plan_create(int n, void** plan);
int ipde_fft_plan_destroy(void* plan);
int ipde_fft_next_size(int n);                 /* smallest even 2^a 3^b 5^c >= n */
int ipde_fft_size_supported(int n);
int ipde_fft_c2c(void* plan, const double* in, double* out, int64_t batch, int sign, double scale, void* stream);
/* out = FFT(m * IFFT(in)), rows of length n: ipde/utilities.py:123-124 (ffourier_multiply) */
int ipde_fft_mul_fft(void* plan, const double* in, double* out, const double* m, int64_t batch, void* stream);

/* ---- periodic background grid ------------------------------------------------------------------ */
int ipde_grid_plan_create(int nx, int ny, double xh, double yh, void** plan);
int ipde_grid_plan_destroy(void* plan);
/* row layout of a half spectrum: frequency index k1 + n1*k2 is stored at row n2*k1 + k2 (n2 == 1: natural) */
int ipde_grid_plan_layout(void* plan, int* n1, int* n2);
/* fc = zeros(nx,ny); fc[phys_lin[i]] = f_compact[i]*step_compact[i]
 * EmbeddedFunction.get_components / get_smoothed_grid_value, ipde/embedded_function.py:135-138,184-194 */
int ipde_grid_scatter_cutoff(void* plan, const double* f_compact, const int64_t* phys_lin,
                             const double* step_compact, int64_t n_phys, double* fc, void* stream);
/* f -= (h^2 sum f) * bumpy : ebdyc.demean_function, ipde/ebdy_collection.py:808-816 */
int ipde_grid_demean(void* plan, double* f, const double* bumpy, void* stream);
/* real (nx,ny) <-> half spectrum (nx, ny/2+1): numpy.fft.fft2 / ifft2(.).real of a real field */
int ipde_grid_fft2_r2c(void* plan, const double* real_in, double* spec, void* stream);
int ipde_grid_fft2_c2r(void* plan, const double* spec_in, double* spec_work, double* real_out, void* stream);
/* spec *= 1/(k^2 + kx^2 + ky^2) (MODHELM) or 1/(-kx^2-ky^2), zero mode -> 0 (POISSON)
 * ipde/solvers/multi_boundary/modified_helmholtz.py:43-46, poisson.py:34-38 */
int ipde_grid_scalar_multiply(void* plan, double* spec, int kind, double k, void* stream);
/* the whole `_grid_solve`: (demean) -> fft2 -> multiplier -> ifft2.real ; uch is kept for the
 * spectral interface interpolation.  modified_helmholtz.py:40-42, poisson.py:30-33 */
int ipde_grid_solve_scalar(void* plan, int kind, double k, double* fc, const double* bumpy, double* uch,
                           double* spec_work, double* uc, void* stream);
/* Stokes `_grid_solve`, ipde/solvers/multi_boundary/stokes.py:34-45: specs of u, v, p are returned too */
int ipde_grid_solve_stokes(void* plan, double* fuc, double* fvc, const double* bumpy, double* uh, double* vh,
                           double* ph, double* spec_work, double* uc, double* vc, double* pc, void* stream);
/* spectra of the grid stress txx = 2u_x - p, txy = u_y + v_x, tyy = 2v_y - p with the spectral derivative of
 * ipde/derivatives.py:25-28 (Nyquist row/column dropped by its .real): ipde/solvers/multi_boundary/vector.py:68-73 */
int ipde_grid_stokes_stress(void* plan, const double* uh, const double* vh, const double* ph, double* txx, double* txy,
                            double* tyy, void* stream);
/* ---- slab-decomposed periodic grid (one slab of rows per rank; ipde_b200/distributed.py).  The same fft2 /
 *      symbol / ifft2 as ipde_grid_solve_scalar, split at the two all-to-all transposes. */
/* real rows (batch, in_pitch doubles) <-> half-complex rows (batch, out_pitch complex); plan: ipde_fft_plan_create(n/2) */
int ipde_fft_r2c_rows(void* plan, const double* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                      void* stream);
int ipde_fft_c2r_rows(void* plan, const double* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                      double scale, void* stream);
/* (rows, P) -> (R, rows, Pc) column blocks, zero padded past P, and back */
int ipde_slab_pack(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, double* out, void* stream);
int ipde_slab_unpack(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, double* out, void* stream);
/* the same for the sub-block [koff, koff + Pk) of every column block: (rows, P) <-> (R, rows, Pk).  Chunks let the
 * all-to-all of one part of the column slab overlap the column pass of another (ipde_b200/distributed.py) */
int ipde_slab_pack_chunk(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, double* out,
                         void* stream);
int ipde_slab_unpack_chunk(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, double* out,
                           void* stream);
/* column FFT + symbol + inverse column FFT on a block of columns [col_offset, col_offset + Pc) of the global
 * half spectrum; plan = ipde_grid_plan_create(nx, 2*(Pc-1), xh, yh) */
int ipde_grid_cols_symbol_slab(void* plan, double* spec_in, double* spec_out, double* spec_copy, int copy_mode, int kind,
                               double k, int ny_global, double yh, int col_offset, void* stream);
/* ---- the same slab solve with the two transposes done by the STORES of the transform kernels, straight into the
 *      peers' buffers over NVLink (CUDA IPC mappings) -- no pack, no all-to-all, no unpack
 *      (ipde_b200/distributed.py, SlabGridSolverP2P).
 * Device buffers that other processes of the box can map: allocate + export a 64-byte handle / map a peer's handle */
int ipde_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle64);
int ipde_ipc_open(const void* handle64, void** dev_ptr);
int ipde_ipc_close(void* dev_ptr);
int ipde_ipc_free(void* dev_ptr);
/* r2c of the `rows` local rows (global rows row0 ...) of ny = 2 (P - 1) reals; entry (row, k) is stored into the column
 * slab (nx, Pc) of rank k / Pc: peer_cols[q] = base of rank q's column slab (R pointers, host array).
 * scratch: (rows, P) complex, local */
int ipde_slab_r2c_rows_scatter(void* row_plan, const double* in, double* scratch, int64_t rows, int64_t in_pitch, int64_t P,
                               int R, int64_t Pc, int64_t row0, void** peer_cols, void* stream);
/* column FFT + symbol + inverse column FFT on this rank's column slab (nx, Pc) (overwritten; spec_work: same size), the
 * last pass storing entry (row i, column col_offset + c) into the row slab (rows, P) of rank i / rows:
 * peer_rows[q] = base of rank q's row slab.  plan = ipde_grid_plan_create(nx, 2*(Pc-1), xh, yh) */
int ipde_slab_cols_symbol_scatter(void* plan, double* spec_in, double* spec_work, int kind, double k, int ny_global,
                                  double yh, int col_offset, int R, int64_t rows, int64_t P, void** peer_rows,
                                  void* stream);
/* 4th-order centred differences, zero in the 2-cell border: ipde/derivatives.py:3-23 */
int ipde_grid_fd4(void* plan, const double* f, double* fx, double* fy, void* stream);
/* periodic cubic interpolation to points: fast_interp.interp2d(k=3) at ipde/ebdy_collection.py:594-603 */
int ipde_grid_interp_cubic(void* plan, const double* f, double x0, double y0, const double* px, const double* py,
                           int64_t npts, double* out, void* stream);
/* out[i] = uc[lin[i]] ; out[pos[i]] += v[i] ; out[pos[i]] = v[i]
 * uc[pna] += gu, radial overwrite, uc*=phys, EmbeddedFunction.load_data: multi_boundary/scalar.py:107-116 */
int ipde_grid_gather(const double* uc, const int64_t* lin, int64_t n, double* out, void* stream);
int ipde_index_add(double* out, const int64_t* pos, const double* v, int64_t n, void* stream);
int ipde_index_set(double* out, const int64_t* pos, const double* v, int64_t n, void* stream);

/* ---- type-2 nonuniform FFT: periodic_interp2d / finufft.nufft2d2(eps=1e-14), scalar.py:80-88,
 *      ebdy_collection.py:604-614.  px_host/py_host: deconvolution factors (host arrays). */
int ipde_nufft2_plan_create(int nx, int ny, double xh, double yh, int fx, int fy, int w, double beta,
                            const double* px_host, const double* py_host, void** plan);
int ipde_nufft2_plan_destroy(void* plan);
/* out_j = Re (1/(nx ny)) sum_k spec[k] (i kx)^dx (i ky)^dy e^{i k.x_j}; (cn1, cn2) = row layout of spec */
int ipde_nufft2_apply(void* plan, const double* spec, int cn1, int cn2, int dx, int dy, const double* x,
                      const double* y, int64_t npts, double* out, void* stream);

/* u, u_x, u_y at the targets from one oversampled transform (the derivatives by differentiating the kernel):
 * the three interpolations of ScalarSolver.__call__, ipde/solvers/multi_boundary/scalar.py:80-88 */
int ipde_nufft2_apply3(void* plan, const double* spec, int cn1, int cn2, const double* x, const double* y,
                       int64_t npts, double* out_v, double* out_dx, double* out_dy, void* stream);

/* ---- radial block -> points: ebdy.interpolate_radial_to_points, ipde/embedded_boundary.py:419-443 */
int ipde_nufft1_plan_create(int n, int fn, int w, double beta, const double* dec_host, void** plan);
int ipde_nufft1_plan_destroy(void* plan);
int ipde_radial_to_points(void* plan, const double* cmat, int M, const double* fr, double* coef_work,
                          const double* theta, const double* t, int64_t npts, double* out,
                          const int64_t* out_pos, void* stream);

/* ---- annular solver: AnnularModifiedHelmholtzSolver / AnnularPoissonSolver,
 *      ipde/annular/modified_helmholtz.py:90-203, ipde/annular/poisson.py:3-21.  All setup arrays host. */
int ipde_annular_plan_create(int n, int M, double k, const double* D01, const double* D12, const double* R01,
                             const double* R12, const double* R02, const double* ibc, const double* obc,
                             const double* ks, const double* psi1, const double* ipsi1, const double* ipsi2,
                             const double* kinv_mode_minor, void** plan);
int ipde_annular_plan_destroy(void* plan);
int ipde_annular_apply(void* plan, const double* in, double* out, void* stream);     /* _apply,  :172-186 */
int ipde_annular_precond(void* plan, const double* in, double* out, void* stream);   /* _optim_preconditioner, :157-159 */
/* solve(RAG, f, ig, og, tol, restart, maxiter): :187-203.  f_sign = -1 for Poisson. iters/relres: host */
int ipde_annular_solve(void* plan, const double* f, const double* ig, const double* og, double f_sign,
                       double tol, int restart, int maxiter, double* u_out, int* iters, double* relres,
                       void* stream);

/* ---- annular Stokes solver: AnnularStokesSolver, ipde/annular/stokes.py:73-541 (mu = viscosity).  Vectors are
 *      (3M-1, n) complex, rows [u_r | u_t | p], mode-minor, Nyquist column zero.  All setup arrays host. */
int ipde_annular_stokes_plan_create(int n, int M, double mu, const double* D01, const double* D12, const double* R01,
                                    const double* R12, const double* R02, const double* P10, const double* ibc,
                                    const double* obc, const double* vi1, const double* ks, const double* psi0,
                                    const double* psi1, const double* ipsi1, const double* ipsi2, const double* combo1,
                                    const double* combo2, const double* m3, const double* kinv_mode_minor, void** plan);
int ipde_annular_stokes_plan_destroy(void* plan);
int ipde_annular_stokes_apply(void* plan, const double* in, double* out, void* stream);     /* _apply_optim_real :321-385 */
int ipde_annular_stokes_precond(void* plan, const double* in, double* out, void* stream);   /* _preconditioner :200-210 */
/* solve(RAG, fr, ft, irg, itg, org, otg, **gmres): :519-541; boundary data pointers may be NULL (zero) */
int ipde_annular_stokes_solve(void* plan, const double* fr, const double* ft, const double* irg, const double* itg,
                              const double* org, const double* otg, double tol, int restart, int maxiter, double* ur,
                              double* ut, double* pr, int* iters, double* relres, void* stream);

/* ---- Stokes (vector) helper glue, radial blocks (M, n), normals/tangents of the boundary curve (n) */
/* inverse 0: (u, v) -> (r, t); 1: (r, t) -> (u, v).  ipde/embedded_boundary.py:521-530 */
int ipde_uv_rt(const double* f0, const double* f1, const double* nx, const double* ny, const double* tx, const double* ty,
               int M, int n, double* out0, double* out1, int inverse, void* stream);
/* traction (x, y) of a radial (Ur, Ut, p) field on the curve picked by the evaluation row `est`; dUr = d/dt Ur;
 * rsp = radial_speed (M, n).  VectorHelper._get_traction_rtp/_uvp, ipde/solvers/internals/vector.py:87-102 */
int ipde_radial_traction(const double* Ur, const double* Ut, const double* dUr, const double* P, const double* D00,
                         const double* est, const double* rsp, const double* nx, const double* ny, const double* tx,
                         const double* ty, int M, int n, double* Tx, double* Ty, void* stream);
/* tau (4n) = sign [rtx - btx | rty - bty | bu | bv], bt = stress . interface normal: vector.py:104-121 */
int ipde_vector_jumps(const double* rtx, const double* rty, const double* txx, const double* txy, const double* tyy,
                      const double* bu, const double* bv, const double* inx, const double* iny, int n, double sign,
                      double* tau, void* stream);
/* pr (M, n) += mean_j [ -(txx_j + tyy_j)/2 + pb_j - sum_i est_i pr[i, j] ]: removes the pressure jump across the
 * interface between the grid side (grid solve + grid-side sources pb) and the annulus side, whose additive constants
 * the stokeslet representations of ipde/solvers/internals/vector.py:104-162 leave free */
int ipde_pressure_continuity(const double* txx, const double* tyy, const double* pb, const double* est, double* pr, int M,
                             int n, void* stream);
/* d/dt of `rows` real rows of length n (plan from ipde_fft_plan_create(n)); work_cplx: rows*n complex.
 * EmbeddedBoundary._radial_grid_t_derivative, ipde/embedded_boundary.py:473-474 */
int ipde_theta_derivative(void* fft_plan, const double* f, double* work_cplx, double* out, int rows, void* stream);

/* ---- layer potentials from point sources: pybie2d *_Layer_Apply / pyfmmlib2d.SFMM,
 *      ipde/solvers/internals/poisson.py:34-35, modified_helmholtz.py:37, stokes.py:26-35 */
/* Chebyshev tables (host) of sqrt(x) e^x K0(x), sqrt(x) e^x K1(x) in t = 4/x - 1 for x >= 2 */
int ipde_set_k0_table(const double* cheb0_host, const double* cheb1_host, int ncheb);
int ipde_eval_k0(const double* x, double* out, int64_t n, void* stream);
int ipde_eval_k1(const double* x, double* out, int64_t n, void* stream);
int ipde_layer_apply_scalar(int kind, double k, const double* sx, const double* sy, const double* q,
                            int64_t nsrc, const double* tx, const double* ty, int64_t ntrg, double* out,
                            int accumulate, void* stream);
/* normal derivative of the same sums at targets with unit normals (tnx, tny): the boundary operator of the homogeneous
 * Neumann correction, examples/multi_modified_helmholtz_neumann.py:117-135 (`s_only` / `s_singular` blocks) */
int ipde_layer_apply_scalar_dn(int kind, double k, const double* sx, const double* sy, const double* q,
                               int64_t nsrc, const double* tx, const double* ty, const double* tnx, const double* tny,
                               int64_t ntrg, double* out, int accumulate, void* stream);
/* setup (QFS): dense out[i*ns + j] = K(t_i, s_j) w_j; layer 0 single, 1 double (normal derivative at the source) */
int ipde_kernel_matrix(int kind, double k, int layer, const double* tx, const double* ty, int64_t nt, const double* sx,
                       const double* sy, const double* snx, const double* sny, const double* w, int64_t ns, double* out,
                       void* stream);
int ipde_layer_apply_stokes(const double* sx, const double* sy, const double* fx, const double* fy,
                            int64_t nsrc, const double* tx, const double* ty, int64_t ntrg, double* out_uvp,
                            int accumulate, void* stream);

/* ---- spectral-Ewald free-space evaluator: the same sums as ipde_layer_apply_scalar in O(S sw^2 + P^2 log P).
 *      Replaces Layer_Apply / Grid_Evaluator at ipde/solvers/multi_boundary/scalar.py:63-71,
 *      ipde/solvers/internals/scalar.py:95-116 and ipde/grid_evaluators/scalar_grid_evaluator.py:188-231,299-307.
 *      Host-side preprocessing (parameters, tables, tiles): ipde_b200/ewald.py.  The FFT box has (px, py) nodes
 *      x0 + i h, y0 + j h; tab*, tile_* are host arrays, sx/sy/targets device arrays (all copied). */
int ipde_ewald_plan_create(int kind, double kappa, double xi, double h, double rc, int m, double x0, double y0, int px,
                           int py, double L, double c0, double c1, int64_t nsrc, const double* sx, const double* sy,
                           int nint, int ndeg, int nlog, const double* tabS_host, const double* tabA_host, int ntiles,
                           const int* tile_ix_host, const int* tile_iy_host, const int* tile_ptr_host,
                           const int* tile_src_host, const int* tile_map_host, void** plan);
int ipde_ewald_plan_destroy(void* plan);
/* on-grid targets: node lin[i] of the box, local-sum slot loc[i] (tile*256 + offset, or -1), added to grid_out[out_pos[i]] */
int ipde_ewald_set_grid_targets(void* plan, int64_t n, const int64_t* lin, const int64_t* loc, const int64_t* out_pos);
int ipde_ewald_set_point_targets(void* plan, int64_t n, const double* tx, const double* ty);
/* q: weighted charges (nsrc).  grid_out[out_pos] += u (if grid targets set and grid_out != NULL);
 * point_out = u or += u at the point targets */
int ipde_ewald_apply(void* plan, const double* q, double* grid_out, double* point_out, int accumulate_points, void* stream);

/* value and gradient with affine weights (stokeslet evaluation, ipde_b200/ewald.py: StokesEvaluator): field c
 * (0 value, 1 d/dx, 2 d/dy) of the potential of q is added to up to two destinations,
 *     dst[.] += (w0 + wx (x - xc) + wy (y - yc)) field_c,
 * gdst_host / pdst_host: host arrays of 6 device addresses [field][2] (0 = unused) for the grid / point targets,
 * gw_host / pw_host: host arrays [field][2][3] of (w0, wx, wy).  Replaces the pyfmmlib2d.SFMM sums of
 * ipde/solvers/internals/stokes.py:26-35 at scale. */
int ipde_ewald_set_gradient_tables(void* plan, const double* dS_host, const double* dA_host);
int ipde_ewald_apply_fields(void* plan, const double* q, double xc, double yc, const double* gw_host,
                            const int64_t* gdst_host, const double* pw_host, const int64_t* pdst_host, void* stream);
/* tests: device pointer/length of a work array: 0 spread S_w, 1 last inverse transform, 2 local sums */
int ipde_ewald_debug_buffer(void* plan, int which, double** ptr, int64_t* n);

/* ---- dense / sparse helpers ----------------------------------------------------------------- */
int ipde_dense_gemv(const double* A, const double* x, double* y, int64_t m, int64_t n, void* stream);
int ipde_lu_solve(const double* LU, const int* perm, const double* b, double* x, int M, void* stream);
/* ipde/sparse_matvec.py:17-40 (zSpMV_viaMKL) and :43-108 (dSpMV_viaMKL) */
int ipde_csr_spmv_d(const int* indptr, const int* indices, const double* data, const double* x, double* y,
                    int64_t nrows, void* stream);
int ipde_csr_spmv_z(const int* indptr, const int* indices, const double* data, const double* x, double* y,
                    int64_t nrows, void* stream);
int ipde_vec_axpby(double a, const double* x, double b, const double* y, double* out, int64_t n, void* stream);
int ipde_vec_mul(const double* x, const double* y, double* out, int64_t n, void* stream);
/* (1 x M) row . (M, n) block: get_boundary_values etc., ipde/solvers/internals/scalar.py:48-67 */
int ipde_rowvec_apply(const double* row, const double* fr, double* out, int M, int64_t n, void* stream);
/* interface jumps of ScalarHelper.__call__, ipde/solvers/internals/scalar.py:73-85: tau = [slp | dlp] */
int ipde_scalar_jumps(const double* bv, const double* bx, const double* by, const double* nx, const double* ny,
                      const double* ur, const double* in_est, int M, int64_t n, double sign, double* tau,
                      void* stream);

/* ---- structured QFS (ipde_b200/qfs_structured.py) ------------------------------------------------
 * Replaces the dense factors of the reference's qfs objects (qfs.Laplace_QFS / Modified_Helmholtz_QFS / Stokes_QFS,
 * applied at ipde/solvers/internals/scalar.py:87-88,106 and internals/vector.py:124-125,150).
 * xhat: (n_in, N) complex, unnormalised N-point FFTs of the densities (FFT order);
 * The operators map real densities to real strengths, Q(-k,-k') = conj Q(k,k'): only output modes k >= 0 are stored.
 * band: (n_out, n_in, 2BH+1, Ks+1) complex, band[o,i,p+BH,k] multiplies mode k-p of input i into mode k of output o;
 * lowT: (n_out, n_in, 2KL+1, KL+1) complex, lowT[o,i,k'+KL,k] (dense block of the lowest modes, transposed);
 * yhat: (ceil(n_out/2), M) complex spectra in FFT order (fully written: modes -k by conjugation, zero beyond Ks),
 *       outputs 2q and 2q+1 packed as real and imaginary part; M >= 2 Ks + 2. */
int ipde_banded_apply_z(const double* xhat, int n_in, int64_t N, const double* band, const double* lowT, int n_out, int BH,
                        int KL, int64_t Ks, double* yhat, int64_t M, void* stream);
int ipde_real_to_complex(const double* x, double* z, int64_t n, void* stream);
int ipde_complex_parts(const double* z, double* re, double* im, int64_t n, void* stream);
/* re[j] (+)= Re z[j*stride], im[j] (+)= Im z[j*stride] (im may be null): sampling of the curve operators' output,
 * Layer_Apply(source -> interface / radial targets) of ipde/solvers/internals/scalar.py:101,113 */
int ipde_complex_parts_strided(const double* z, double* re, double* im, int64_t n, int64_t stride, int accumulate,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif
