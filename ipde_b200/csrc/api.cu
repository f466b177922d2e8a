// extern "C" surface of libipde_b200.so (declared in include/ipde_b200.h).
#include "../../include/ipde_b200.h"
#include "annular.cuh"
#include "dense.cuh"
#include "ewald.cuh"
#include "fft.cuh"
#include "grid.cuh"
#include "layer.cuh"
#include "nufft.cuh"
#include "radial.cuh"
#include "vector.cuh"
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <mutex>

static thread_local char g_err[1024] = "";

void ipde_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void ipde_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-kernel profile: one event after every launch, elapsed time between consecutive events
#include <map>
#include <string>
#include <vector>
struct ProfRec { std::string name; double bytes; cudaEvent_t ev; };
static bool g_prof_on = false;
static cudaStream_t g_prof_stream = 0;
static cudaEvent_t g_prof_begin = nullptr;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static const char* g_prof_name = nullptr;
static double g_prof_bytes = 0.0;
void ipde_prof_tag(const char* name, double bytes) { g_prof_name = name; g_prof_bytes = bytes; }
void ipde_prof_mark(const char* file, int line) {
    if (!g_prof_on) { g_prof_name = nullptr; return; }
    ProfRec r;
    if (g_prof_name) { r.name = g_prof_name; r.bytes = g_prof_bytes; }
    else {
        const char* base = strrchr(file, '/');
        r.name = std::string(base ? base + 1 : file) + ":" + std::to_string(line);
        r.bytes = 0.0;
    }
    g_prof_name = nullptr;
    if (!g_prof_pool.empty()) { r.ev = g_prof_pool.back(); g_prof_pool.pop_back(); }
    else if (cudaEventCreate(&r.ev) != cudaSuccess) return;
    cudaEventRecord(r.ev, g_prof_stream);
    g_prof_recs.push_back(r);
}

#define ST(s) ((cudaStream_t)(s))
#define CP(p) (reinterpret_cast<const cplx*>(p))
#define MP(p) (reinterpret_cast<cplx*>(p))

extern "C" {

const char* ipde_last_error(void) { return g_err; }
int ipde_version(void) { return 100; }
int ipde_device_sm_count(void) { return ipde_num_sms(); }
int64_t ipde_launch_count(void) { return (int64_t)g_launches.load(); }

int ipde_profile_start(void* stream) {
    g_prof_stream = ST(stream);
    for (auto& r : g_prof_recs) g_prof_pool.push_back(r.ev);
    g_prof_recs.clear();
    if (!g_prof_begin) CUDA_TRY(cudaEventCreate(&g_prof_begin));
    CUDA_TRY(cudaEventRecord(g_prof_begin, g_prof_stream));
    g_prof_on = true;
    return IPDE_OK;
}

// Writes one line per kernel name: "name count total_ms total_bytes\n" (sorted by time) into `out`.
int ipde_profile_stop(char* out, int64_t cap) {
    g_prof_on = false;
    CUDA_TRY(cudaStreamSynchronize(g_prof_stream));
    struct Acc { int64_t n = 0; double ms = 0.0, bytes = 0.0; };
    std::map<std::string, Acc> acc;
    cudaEvent_t prev = g_prof_begin;
    for (auto& r : g_prof_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, prev, r.ev);
        Acc& a = acc[r.name];
        a.n += 1; a.ms += ms; a.bytes += r.bytes;
        prev = r.ev;
    }
    std::vector<std::pair<std::string, Acc>> v(acc.begin(), acc.end());
    std::sort(v.begin(), v.end(), [](const std::pair<std::string, Acc>& a, const std::pair<std::string, Acc>& b) { return a.second.ms > b.second.ms; });
    std::string text;
    for (auto& kv : v) {
        char line[512];
        snprintf(line, sizeof(line), "%s %lld %.6f %.0f\n", kv.first.c_str(), (long long)kv.second.n, kv.second.ms, kv.second.bytes);
        text += line;
    }
    if (out && cap > 0) {
        const size_t nmax = (size_t)cap - 1;
        const size_t ncopy = text.size() < nmax ? text.size() : nmax;
        memcpy(out, text.data(), ncopy);
        out[ncopy] = 0;
    }
    for (auto& r : g_prof_recs) g_prof_pool.push_back(r.ev);
    g_prof_recs.clear();
    return IPDE_OK;
}

// ---- FFT
int ipde_fft_plan_create(int n, void** plan) {
    FftPlan* p = nullptr;
    int rc = fft_plan_create(n, &p);
    *plan = p;
    return rc;
}
int ipde_fft_plan_destroy(void* plan) { fft_plan_destroy((FftPlan*)plan); return IPDE_OK; }
int ipde_fft_next_size(int n) { return fft_next_size(n); }
int ipde_fft_size_supported(int n) { return fft_size_supported(n) ? 1 : 0; }
int ipde_fft_c2c(void* plan, const double* in, double* out, int64_t batch, int sign, double scale, void* stream) {
    FftPlan* p = (FftPlan*)plan;
    return fft_rows(p, CP(in), MP(out), batch, p->d.n, p->d.n, sign, scale, ST(stream));
}
int ipde_fft_mul_fft(void* plan, const double* in, double* out, const double* m, int64_t batch, void* stream) {
    return fft_mul_fft_rows((FftPlan*)plan, CP(in), MP(out), m, batch, ST(stream));
}

// ---- grid
int ipde_grid_plan_create(int nx, int ny, double xh, double yh, void** plan) {
    GridPlan* g = nullptr;
    int rc = grid_plan_create(nx, ny, xh, yh, &g);
    *plan = g;
    return rc;
}
int ipde_grid_plan_destroy(void* plan) { grid_plan_destroy((GridPlan*)plan); return IPDE_OK; }
int ipde_grid_plan_layout(void* plan, int* n1, int* n2) {
    GridPlan* g = (GridPlan*)plan;
    *n1 = g->n1; *n2 = g->n2;
    return IPDE_OK;
}
int ipde_grid_scatter_cutoff(void* plan, const double* f_compact, const int64_t* phys_lin, const double* step_compact,
                             int64_t n_phys, double* fc, void* stream) {
    return grid_scatter_cutoff((GridPlan*)plan, f_compact, phys_lin, step_compact, n_phys, fc, ST(stream));
}
int ipde_grid_demean(void* plan, double* f, const double* bumpy, void* stream) {
    return grid_demean((GridPlan*)plan, f, bumpy, ST(stream));
}
int ipde_grid_fft2_r2c(void* plan, const double* real_in, double* spec, void* stream) {
    return grid_fft2_forward((GridPlan*)plan, real_in, MP(spec), ST(stream));
}
int ipde_grid_fft2_c2r(void* plan, const double* spec_in, double* spec_work, double* real_out, void* stream) {
    return grid_fft2_inverse((GridPlan*)plan, CP(spec_in), MP(spec_work), real_out, ST(stream));
}
int ipde_grid_scalar_multiply(void* plan, double* spec, int kind, double k, void* stream) {
    return grid_scalar_multiply((GridPlan*)plan, MP(spec), kind, k, ST(stream));
}
int ipde_grid_solve_scalar(void* plan, int kind, double k, double* fc, const double* bumpy, double* uch, double* spec_work,
                           double* uc, void* stream) {
    GridPlan* g = (GridPlan*)plan;
    int rc = IPDE_OK;
    if (kind == IPDE_KIND_POISSON) {
        ARG_CHECK(bumpy != nullptr, "Poisson grid solve needs the bump array (ebdyc.ready_bump)");
        rc = grid_demean(g, fc, bumpy, ST(stream));
        if (rc) return rc;
    }
    return grid_solve_scalar_fused(g, kind, k, fc, MP(uch), MP(spec_work), uc, ST(stream));
}
int ipde_grid_solve_stokes(void* plan, double* fuc, double* fvc, const double* bumpy, double* uh, double* vh, double* ph,
                           double* spec_work, double* uc, double* vc, double* pc, void* stream) {
    GridPlan* g = (GridPlan*)plan;
    ARG_CHECK(bumpy != nullptr, "Stokes grid solve needs the bump array (ebdyc.ready_bump)");
    int rc = grid_demean(g, fuc, bumpy, ST(stream));
    if (!rc) rc = grid_demean(g, fvc, bumpy, ST(stream));
    if (!rc) rc = grid_fft2_forward(g, fuc, MP(uh), ST(stream));
    if (!rc) rc = grid_fft2_forward(g, fvc, MP(vh), ST(stream));
    if (!rc) rc = grid_stokes_multiply(g, MP(uh), MP(vh), MP(ph), ST(stream));
    if (!rc) rc = grid_fft2_inverse(g, CP(uh), MP(spec_work), uc, ST(stream));
    if (!rc) rc = grid_fft2_inverse(g, CP(vh), MP(spec_work), vc, ST(stream));
    if (!rc) rc = grid_fft2_inverse(g, CP(ph), MP(spec_work), pc, ST(stream));
    return rc;
}
int ipde_grid_stokes_stress(void* plan, const double* uh, const double* vh, const double* ph, double* txx, double* txy,
                            double* tyy, void* stream) {
    return grid_stokes_stress((GridPlan*)plan, CP(uh), CP(vh), CP(ph), MP(txx), MP(txy), MP(tyy), ST(stream));
}
// ---- slab-decomposed grid (ipde_b200/distributed.py)
int ipde_fft_r2c_rows(void* plan, const double* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, void* stream) {
    return fft_r2c_rows((FftPlan*)plan, in, MP(out), batch, in_pitch, out_pitch, ST(stream));
}
int ipde_fft_c2r_rows(void* plan, const double* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                      double scale, void* stream) {
    return fft_c2r_rows((FftPlan*)plan, CP(in), out, batch, in_pitch, out_pitch, scale, ST(stream));
}
int ipde_slab_pack(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, double* out, void* stream) {
    return grid_slab_pack(CP(in), rows, P, R, Pc, 0, Pc, MP(out), ST(stream));
}
int ipde_slab_unpack(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, double* out, void* stream) {
    return grid_slab_unpack(CP(in), rows, P, R, Pc, 0, Pc, MP(out), ST(stream));
}
int ipde_slab_pack_chunk(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, double* out,
                         void* stream) {
    return grid_slab_pack(CP(in), rows, P, R, Pc, koff, Pk, MP(out), ST(stream));
}
int ipde_slab_unpack_chunk(const double* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, double* out,
                           void* stream) {
    return grid_slab_unpack(CP(in), rows, P, R, Pc, koff, Pk, MP(out), ST(stream));
}
// device buffers shared between the processes of one box (CUDA IPC)
int ipde_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle64) {
    ARG_CHECK(bytes > 0 && dev_ptr && handle64, "ipc alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
    CUDA_TRY(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        ipde_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return IPDE_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return IPDE_OK;
}
int ipde_ipc_open(const void* handle64, void** dev_ptr) {
    ARG_CHECK(handle64 && dev_ptr, "ipc open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return IPDE_OK;
}
int ipde_ipc_close(void* dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return IPDE_OK;
}
int ipde_ipc_free(void* dev_ptr) {
    if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
    return IPDE_OK;
}
static int fill_scatter(SlabScatter& S, int R, void** peers) {
    ARG_CHECK(R >= 1 && R <= IPDE_SLAB_MAX_RANKS && peers, "slab scatter: 1..8 ranks");
    S = slab_scatter_none();
    S.R = R;
    for (int q = 0; q < R; ++q) {
        ARG_CHECK(peers[q], "slab scatter: null peer buffer");
        S.base[q] = (cplx*)peers[q];
    }
    return IPDE_OK;
}
int ipde_slab_r2c_rows_scatter(void* row_plan, const double* in, double* scratch, int64_t rows, int64_t in_pitch, int64_t P,
                               int R, int64_t Pc, int64_t row0, void** peer_cols, void* stream) {
    SlabScatter S;
    int rc = fill_scatter(S, R, peer_cols);
    if (rc) return rc;
    S.rows = (int)rows; S.Pc = Pc; S.P = P; S.row0 = row0; S.col0 = 0;
    return fft_r2c_rows_scatter((FftPlan*)row_plan, in, MP(scratch), rows, in_pitch, P, S, ST(stream));
}
int ipde_slab_cols_symbol_scatter(void* plan, double* spec_in, double* spec_work, int kind, double k, int ny_global,
                                  double yh, int col_offset, int R, int64_t rows, int64_t P, void** peer_rows,
                                  void* stream) {
    SlabScatter S;
    int rc = fill_scatter(S, R, peer_rows);
    if (rc) return rc;
    GridPlan* g = (GridPlan*)plan;
    S.rows = (int)rows; S.Pc = g->pitch; S.P = P; S.row0 = 0; S.col0 = col_offset;
    return grid_cols_symbol_slab_scatter(g, MP(spec_in), MP(spec_work), kind, k, ny_global, yh, col_offset, S, ST(stream));
}
int ipde_grid_cols_symbol_slab(void* plan, double* spec_in, double* spec_out, double* spec_copy, int copy_mode, int kind,
                               double k, int ny_global, double yh, int col_offset, void* stream) {
    return grid_cols_symbol_slab((GridPlan*)plan, MP(spec_in), MP(spec_out), MP(spec_copy), copy_mode, kind, k, ny_global, yh,
                                 col_offset, ST(stream));
}
int ipde_grid_fd4(void* plan, const double* f, double* fx, double* fy, void* stream) {
    return grid_fd4((GridPlan*)plan, f, fx, fy, ST(stream));
}
int ipde_grid_interp_cubic(void* plan, const double* f, double x0, double y0, const double* px, const double* py,
                           int64_t npts, double* out, void* stream) {
    return grid_interp_cubic((GridPlan*)plan, f, x0, y0, px, py, npts, out, ST(stream));
}
int ipde_grid_gather(const double* uc, const int64_t* lin, int64_t n, double* out, void* stream) {
    return grid_gather(uc, lin, n, out, ST(stream));
}
int ipde_index_add(double* out, const int64_t* pos, const double* v, int64_t n, void* stream) {
    return grid_index_add(out, pos, v, n, ST(stream));
}
int ipde_index_set(double* out, const int64_t* pos, const double* v, int64_t n, void* stream) {
    return grid_index_set(out, pos, v, n, ST(stream));
}

// ---- NUFFT
int ipde_nufft2_plan_create(int nx, int ny, double xh, double yh, int fx, int fy, int w, double beta,
                            const double* px_host, const double* py_host, void** plan) {
    Nufft2Plan* p = nullptr;
    int rc = nufft2_plan_create(nx, ny, xh, yh, fx, fy, w, beta, px_host, py_host, &p);
    *plan = p;
    return rc;
}
int ipde_nufft2_plan_destroy(void* plan) { nufft2_plan_destroy((Nufft2Plan*)plan); return IPDE_OK; }
int ipde_nufft2_apply3(void* plan, const double* spec, int cn1, int cn2, const double* x, const double* y, int64_t npts,
                       double* out_v, double* out_dx, double* out_dy, void* stream) {
    return nufft2_apply3((Nufft2Plan*)plan, CP(spec), cn1, cn2, x, y, npts, out_v, out_dx, out_dy, ST(stream));
}
int ipde_nufft2_apply(void* plan, const double* spec, int cn1, int cn2, int dx, int dy, const double* x, const double* y,
                      int64_t npts, double* out, void* stream) {
    return nufft2_apply((Nufft2Plan*)plan, CP(spec), cn1, cn2, dx, dy, x, y, npts, out, ST(stream));
}
int ipde_nufft1_plan_create(int n, int fn, int w, double beta, const double* dec_host, void** plan) {
    Nufft1Plan* p = nullptr;
    int rc = nufft1_plan_create(n, fn, w, beta, dec_host, &p);
    *plan = p;
    return rc;
}
int ipde_nufft1_plan_destroy(void* plan) { nufft1_plan_destroy((Nufft1Plan*)plan); return IPDE_OK; }
int ipde_radial_to_points(void* plan, const double* cmat, int M, const double* fr, double* coef_work, const double* theta,
                          const double* t, int64_t npts, double* out, const int64_t* out_pos, void* stream) {
    return radial_to_points((Nufft1Plan*)plan, cmat, M, fr, coef_work, theta, t, npts, out, out_pos, ST(stream));
}

// ---- annular
int ipde_annular_plan_create(int n, int M, double k, const double* D01, const double* D12, const double* R01,
                             const double* R12, const double* R02, const double* ibc, const double* obc, const double* ks,
                             const double* psi1, const double* ipsi1, const double* ipsi2, const double* kinv_mode_minor,
                             void** plan) {
    AnnularPlan* p = nullptr;
    int rc = annular_plan_create(n, M, k, D01, D12, R01, R12, R02, ibc, obc, ks, psi1, ipsi1, ipsi2, kinv_mode_minor, &p);
    *plan = p;
    return rc;
}
int ipde_annular_plan_destroy(void* plan) { annular_plan_destroy((AnnularPlan*)plan); return IPDE_OK; }
int ipde_annular_apply(void* plan, const double* in, double* out, void* stream) {
    return annular_apply((AnnularPlan*)plan, CP(in), MP(out), ST(stream));
}
int ipde_annular_precond(void* plan, const double* in, double* out, void* stream) {
    return annular_precond((AnnularPlan*)plan, CP(in), MP(out), ST(stream));
}
int ipde_annular_solve(void* plan, const double* f, const double* ig, const double* og, double f_sign, double tol,
                       int restart, int maxiter, double* u_out, int* iters, double* relres, void* stream) {
    return annular_solve((AnnularPlan*)plan, f, ig, og, f_sign, tol, restart, maxiter, u_out, iters, relres, ST(stream));
}

// ---- annular Stokes
int ipde_annular_stokes_plan_create(int n, int M, double mu, const double* D01, const double* D12, const double* R01,
                                    const double* R12, const double* R02, const double* P10, const double* ibc,
                                    const double* obc, const double* vi1, const double* ks, const double* psi0,
                                    const double* psi1, const double* ipsi1, const double* ipsi2, const double* combo1,
                                    const double* combo2, const double* m3, const double* kinv_mode_minor, void** plan) {
    AnnularStokesPlan* p = nullptr;
    int rc = annular_stokes_plan_create(n, M, mu, D01, D12, R01, R12, R02, P10, ibc, obc, vi1, ks, psi0, psi1, ipsi1, ipsi2,
                                        combo1, combo2, m3, kinv_mode_minor, &p);
    *plan = p;
    return rc;
}
int ipde_annular_stokes_plan_destroy(void* plan) { annular_stokes_plan_destroy((AnnularStokesPlan*)plan); return IPDE_OK; }
int ipde_annular_stokes_apply(void* plan, const double* in, double* out, void* stream) {
    return annular_stokes_apply((AnnularStokesPlan*)plan, CP(in), MP(out), ST(stream));
}
int ipde_annular_stokes_precond(void* plan, const double* in, double* out, void* stream) {
    return annular_stokes_precond((AnnularStokesPlan*)plan, CP(in), MP(out), ST(stream));
}
int ipde_annular_stokes_solve(void* plan, const double* fr, const double* ft, const double* irg, const double* itg,
                              const double* org, const double* otg, double tol, int restart, int maxiter, double* ur,
                              double* ut, double* pr, int* iters, double* relres, void* stream) {
    return annular_stokes_solve((AnnularStokesPlan*)plan, fr, ft, irg, itg, org, otg, tol, restart, maxiter, ur, ut, pr, iters,
                                relres, ST(stream));
}

// ---- Stokes helper glue
int ipde_uv_rt(const double* f0, const double* f1, const double* nx, const double* ny, const double* tx, const double* ty, int M,
               int n, double* out0, double* out1, int inverse, void* stream) {
    return vec_uv_rt(f0, f1, nx, ny, tx, ty, M, n, out0, out1, inverse, ST(stream));
}
int ipde_radial_traction(const double* Ur, const double* Ut, const double* dUr, const double* P, const double* D00,
                         const double* est, const double* rsp, const double* nx, const double* ny, const double* tx,
                         const double* ty, int M, int n, double* Tx, double* Ty, void* stream) {
    return vec_radial_traction(Ur, Ut, dUr, P, D00, est, rsp, nx, ny, tx, ty, M, n, Tx, Ty, ST(stream));
}
int ipde_vector_jumps(const double* rtx, const double* rty, const double* txx, const double* txy, const double* tyy,
                      const double* bu, const double* bv, const double* inx, const double* iny, int n, double sign,
                      double* tau, void* stream) {
    return vec_jumps(rtx, rty, txx, txy, tyy, bu, bv, inx, iny, n, sign, tau, ST(stream));
}
int ipde_pressure_continuity(const double* txx, const double* tyy, const double* pb, const double* est, double* pr, int M,
                             int n, void* stream) {
    return vec_pressure_continuity(txx, tyy, pb, est, pr, M, n, ST(stream));
}
int ipde_theta_derivative(void* fft_plan, const double* f, double* work_cplx, double* out, int rows, void* stream) {
    return vec_theta_derivative((FftPlan*)fft_plan, f, work_cplx, out, rows, ST(stream));
}

// ---- spectral-Ewald free-space evaluator
int ipde_ewald_plan_create(int kind, double kappa, double xi, double h, double rc, int m, double x0, double y0, int px,
                           int py, double L, double c0, double c1, int64_t nsrc, const double* sx, const double* sy,
                           int nint, int ndeg, int nlog, const double* tabS_host, const double* tabA_host, int ntiles,
                           const int* tile_ix_host, const int* tile_iy_host, const int* tile_ptr_host,
                           const int* tile_src_host, const int* tile_map_host, void** plan) {
    EwaldPlan* p = nullptr;
    int rc_ = ewald_plan_create(kind, kappa, xi, h, rc, m, x0, y0, px, py, L, c0, c1, nsrc, sx, sy, nint, ndeg, nlog,
                                tabS_host, tabA_host, ntiles, tile_ix_host, tile_iy_host, tile_ptr_host, tile_src_host,
                                tile_map_host, &p);
    *plan = p;
    return rc_;
}
int ipde_ewald_plan_destroy(void* plan) { ewald_plan_destroy((EwaldPlan*)plan); return IPDE_OK; }
int ipde_ewald_set_grid_targets(void* plan, int64_t n, const int64_t* lin, const int64_t* loc, const int64_t* out_pos) {
    return ewald_set_grid_targets((EwaldPlan*)plan, n, lin, loc, out_pos);
}
int ipde_ewald_set_point_targets(void* plan, int64_t n, const double* tx, const double* ty) {
    return ewald_set_point_targets((EwaldPlan*)plan, n, tx, ty);
}
int ipde_ewald_apply(void* plan, const double* q, double* grid_out, double* point_out, int accumulate_points, void* stream) {
    return ewald_apply((EwaldPlan*)plan, q, grid_out, point_out, accumulate_points, ST(stream));
}

int ipde_ewald_set_gradient_tables(void* plan, const double* dS_host, const double* dA_host) {
    return ewald_set_gradient_tables((EwaldPlan*)plan, dS_host, dA_host);
}
int ipde_ewald_apply_fields(void* plan, const double* q, double xc, double yc, const double* gw_host, const int64_t* gdst_host,
                            const double* pw_host, const int64_t* pdst_host, void* stream) {
    return ewald_apply_fields((EwaldPlan*)plan, q, xc, yc, gw_host, reinterpret_cast<double* const*>(gdst_host), pw_host,
                              reinterpret_cast<double* const*>(pdst_host), ST(stream));
}
int ipde_ewald_debug_buffer(void* plan, int which, double** ptr, int64_t* n) {
    return ewald_debug_buffer((EwaldPlan*)plan, which, ptr, n);
}

// ---- layer potentials
int ipde_set_k0_table(const double* cheb0_host, const double* cheb1_host, int ncheb) { return layer_set_k0_table(cheb0_host, cheb1_host, ncheb); }
int ipde_eval_k0(const double* x, double* out, int64_t n, void* stream) { return layer_eval_k0(x, out, n, 0, ST(stream)); }
int ipde_eval_k1(const double* x, double* out, int64_t n, void* stream) { return layer_eval_k0(x, out, n, 1, ST(stream)); }
int ipde_layer_apply_scalar(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                            const double* tx, const double* ty, int64_t ntrg, double* out, int accumulate, void* stream) {
    return layer_apply_scalar(kind, k, sx, sy, q, nsrc, tx, ty, ntrg, out, accumulate, ST(stream));
}
int ipde_layer_apply_scalar_dn(int kind, double k, const double* sx, const double* sy, const double* q, int64_t nsrc,
                               const double* tx, const double* ty, const double* tnx, const double* tny, int64_t ntrg,
                               double* out, int accumulate, void* stream) {
    return layer_apply_scalar_dn(kind, k, sx, sy, q, nsrc, tx, ty, tnx, tny, ntrg, out, accumulate, ST(stream));
}
int ipde_kernel_matrix(int kind, double k, int layer, const double* tx, const double* ty, int64_t nt, const double* sx,
                       const double* sy, const double* snx, const double* sny, const double* w, int64_t ns, double* out,
                       void* stream) {
    return layer_kernel_matrix(kind, k, layer, tx, ty, nt, sx, sy, snx, sny, w, ns, out, ST(stream));
}
int ipde_layer_apply_stokes(const double* sx, const double* sy, const double* fx, const double* fy, int64_t nsrc,
                            const double* tx, const double* ty, int64_t ntrg, double* out_uvp, int accumulate, void* stream) {
    return layer_apply_stokes(sx, sy, fx, fy, nsrc, tx, ty, ntrg, out_uvp, accumulate, ST(stream));
}

// ---- dense / sparse
int ipde_dense_gemv(const double* A, const double* x, double* y, int64_t m, int64_t n, void* stream) {
    return dense_gemv(A, x, y, m, n, ST(stream));
}
int ipde_lu_solve(const double* LU, const int* perm, const double* b, double* x, int M, void* stream) {
    return dense_lu_solve(LU, perm, b, x, M, ST(stream));
}
int ipde_csr_spmv_d(const int* indptr, const int* indices, const double* data, const double* x, double* y, int64_t nrows,
                    void* stream) {
    return csr_spmv_d(indptr, indices, data, x, y, nrows, ST(stream));
}
int ipde_csr_spmv_z(const int* indptr, const int* indices, const double* data, const double* x, double* y, int64_t nrows,
                    void* stream) {
    return csr_spmv_z(indptr, indices, CP(data), CP(x), MP(y), nrows, ST(stream));
}
int ipde_vec_axpby(double a, const double* x, double b, const double* y, double* out, int64_t n, void* stream) {
    return vec_axpby(a, x, b, y, out, n, ST(stream));
}
int ipde_vec_mul(const double* x, const double* y, double* out, int64_t n, void* stream) {
    return vec_mul(x, y, out, n, ST(stream));
}
int ipde_rowvec_apply(const double* row, const double* fr, double* out, int M, int64_t n, void* stream) {
    return rowvec_apply(row, fr, out, M, n, ST(stream));
}
int ipde_scalar_jumps(const double* bv, const double* bx, const double* by, const double* nx, const double* ny,
                      const double* ur, const double* in_est, int M, int64_t n, double sign, double* tau, void* stream) {
    return scalar_jumps(bv, bx, by, nx, ny, ur, in_est, M, n, sign, tau, ST(stream));
}
int ipde_banded_apply_z(const double* xhat, int n_in, int64_t N, const double* band, const double* lowT, int n_out, int BH,
                        int KL, int64_t Ks, double* yhat, int64_t M, void* stream) {
    return banded_apply_z(CP(xhat), n_in, N, CP(band), CP(lowT), n_out, BH, KL, Ks, MP(yhat), M, ST(stream));
}
int ipde_real_to_complex(const double* x, double* z, int64_t n, void* stream) {
    return real_to_complex(x, MP(z), n, ST(stream));
}
int ipde_complex_parts(const double* z, double* re, double* im, int64_t n, void* stream) {
    return complex_parts(CP(z), re, im, n, ST(stream));
}
int ipde_complex_parts_strided(const double* z, double* re, double* im, int64_t n, int64_t stride, int accumulate,
                               void* stream) {
    return complex_parts_strided(CP(z), re, im, n, stride, accumulate, ST(stream));
}

}  // extern "C"
