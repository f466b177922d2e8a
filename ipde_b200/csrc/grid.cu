// Periodic background-grid kernels: compact->grid scatter with cutoff, demean, 2-D real FFT,
// spectral multipliers (Poisson / modified Helmholtz / Stokes), finite differences, finalize.
//
// Replaces, on the device:
//   EmbeddedFunction.get_components / get_smoothed_grid_value   ipde/embedded_function.py:135-138,184-194
//   ebdyc.demean_function / grid_integral                       ipde/ebdy_collection.py:808-816
//   ModifiedHelmholtzSolver._grid_solve                         ipde/solvers/multi_boundary/modified_helmholtz.py:40-46
//   PoissonSolver._grid_solve                                   ipde/solvers/multi_boundary/poisson.py:30-38
//   StokesSolver._grid_solve                                    ipde/solvers/multi_boundary/stokes.py:34-50
//   derivatives.fd_x_4 / fd_y_4                                 ipde/derivatives.py:3-23
//   uc[pna] += gu ; uc *= phys ; EmbeddedFunction.load_data     ipde/solvers/multi_boundary/scalar.py:107-116
#include "grid.cuh"
#include <algorithm>

static const double TWO_PI = 6.283185307179586476925286766559;

// ------------------------------------------------------------------------------------------
int grid_plan_create(int nx, int ny, double xh, double yh, GridPlan** out) {
    ARG_CHECK(nx >= 4 && ny >= 4 && ny % 2 == 0, "grid sizes must be >= 4 and ny even");
    GridPlan* g = new GridPlan();
    g->nx = nx; g->ny = ny; g->xh = xh; g->yh = yh;
    g->pitch = ny / 2 + 1;
    int rc = fft_plan_create(ny / 2, &g->row);
    if (rc) { delete g; return rc; }
    // columns: one pass if at least 4 adjacent columns fit next to each other in shared memory,
    // otherwise a four-step split nx = n1*n2
    // (power-of-two lengths from 1024 are always split: both factors are then <= 64 up to nx = 4096 and run in the
    // register-resident short-line kernels of fft_short.cuh, three streaming passes that beat one shared-memory
    // pass with a single resident CTA per SM)
    const bool split_pow2 = fft_short_enabled() && nx >= 1024 && (nx & (nx - 1)) == 0;
    if ((int64_t)nx * 4 <= IPDE_FFT_SMEM_MAX && !split_pow2) {
        g->n1 = nx; g->n2 = 1;
        rc = fft_plan_create(nx, &g->colA);
        if (rc) { grid_plan_destroy(g); return rc; }
    } else {
        int best = 1;
        for (int d = 1; (int64_t)d * d <= nx; ++d)
            if (nx % d == 0 && fft_size_supported(d) && fft_size_supported(nx / d)) best = d;
        g->n2 = best; g->n1 = nx / best;
        if (best == 1 || (int64_t)g->n1 * 4 > IPDE_FFT_SMEM_MAX) {
            // no usable split: long-line global path
            g->n1 = nx; g->n2 = 1;
            rc = fft_plan_create(nx, &g->colA);
            if (rc) { grid_plan_destroy(g); return rc; }
        } else {
            rc = fft_plan_create(g->n1, &g->colA);
            if (!rc) rc = fft_plan_create(g->n2, &g->colB);
            if (rc) { grid_plan_destroy(g); return rc; }
            std::vector<cplx> tw(nx);
            const long double tp = 6.283185307179586476925286766559005768L;
            for (int j = 0; j < nx; ++j) {
                long double a = -tp * (long double)j / (long double)nx;
                tw[j] = make_double2((double)cosl(a), (double)sinl(a));
            }
            CUDA_TRY(cudaMalloc(&g->tw_big, sizeof(cplx) * (size_t)nx));
            CUDA_TRY(cudaMemcpy(g->tw_big, tw.data(), sizeof(cplx) * (size_t)nx, cudaMemcpyHostToDevice));
        }
    }
    *out = g;
    return IPDE_OK;
}

void grid_plan_destroy(GridPlan* g) {
    if (!g) return;
    fft_plan_destroy(g->row);
    fft_plan_destroy(g->colA);
    fft_plan_destroy(g->colB);
    cudaFree(g->tw_big);
    cudaFree(g->reduce_buf);
    delete g;
}

// forward column transform of the half spectrum (in place or out of place)
int grid_cols_forward(GridPlan* g, const cplx* in, cplx* out, cudaStream_t st) {
    const int64_t P = g->pitch;
    if (g->n2 == 1) return fft_cols(g->colA, in, out, P, P, 1, 0, -1, 1.0, 0, nullptr, 0, st);
    // x = n2*a + b.  A: FFT over a for every b (+ twiddle), B: FFT over b for every k1.
    int rc = fft_cols(g->colA, in, out, (int64_t)g->n2 * P, P, g->n2, P, -1, 1.0, 1, g->tw_big, g->nx, st);
    if (rc) return rc;
    return fft_cols(g->colB, out, out, P, P, g->n1, (int64_t)g->n2 * P, -1, 1.0, 0, nullptr, 0, st);
}

int grid_cols_inverse(GridPlan* g, const cplx* in, cplx* out, cudaStream_t st) {
    return grid_cols_inverse_n(g, in, out, g->pitch, st);
}

// the same for the first `ncols` columns only (zero-padded spectra: the columns beyond hold zeros and stay untouched)
int grid_cols_inverse_n(GridPlan* g, const cplx* in, cplx* out, int64_t ncols, cudaStream_t st) {
    const int64_t P = g->pitch;
    if (g->n2 == 1) return fft_cols(g->colA, in, out, P, ncols, 1, 0, +1, 1.0, 0, nullptr, 0, st);
    int rc = fft_cols(g->colB, in, out, P, ncols, g->n1, (int64_t)g->n2 * P, +1, 1.0, 0, nullptr, 0, st);
    if (rc) return rc;
    return fft_cols(g->colA, out, out, (int64_t)g->n2 * P, ncols, g->n2, P, +1, 1.0, 2, g->tw_big, g->nx, st);
}

int grid_fft2_forward(GridPlan* g, const double* real_in, cplx* spec, cudaStream_t st) {
    int rc = fft_r2c_rows(g->row, real_in, spec, g->nx, g->ny, g->pitch, st);
    if (rc) return rc;
    return grid_cols_forward(g, spec, spec, st);
}

int grid_fft2_inverse(GridPlan* g, const cplx* spec_in, cplx* spec_work, double* real_out, cudaStream_t st) {
    int rc = grid_cols_inverse(g, spec_in, spec_work, st);
    if (rc) return rc;
    return fft_c2r_rows(g->row, spec_work, real_out, g->nx, g->pitch, g->ny, 1.0 / ((double)g->nx * (double)g->ny) * 2.0, st);
}

// ------------------------------------------------------------------------------------------
// Fused spectral pass over the columns of a half spectrum:
//     load tile -> [forward FFT] -> multiply by a real symbol -> inverse FFT -> store
// with an optional copy of the spectrum (before or after the multiplication) to a second array.  It is the
// whole column direction of a grid whose columns fit one tile, and the inner ("B") pass of the four-step
// split otherwise; either way the multiplier costs no pass over memory of its own.
//   symbol kinds: IPDE_KIND_POISSON / IPDE_KIND_MODHELM analytic (reference multi_boundary/poisson.py:34-38,
//   modified_helmholtz.py:43-46), 3: table with the layout of the spectrum (free-space evaluator, ewald.cu)
__device__ __forceinline__ cplx colmult_apply(const ColMult& M, cplx v, int64_t addr, int o, int e, int col);

template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_cols_mult(FftPlanDev P, const cplx* in, cplx* out, cplx* copy, int64_t elem_stride,
            int64_t n_inner, int64_t outer_stride, int C, int logC, int do_fwd, int copy_mode, ColMult M, RowWindow W) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t tiles = (n_inner + C - 1) >> logC;
    const int o = (int)(blockIdx.x / tiles);
    const int64_t c0 = (blockIdx.x - (int64_t)o * tiles) << logC;
    const int cw = (int)min((int64_t)C, n_inner - c0);
    const int64_t base = (int64_t)o * outer_stride + c0;
    const int total = n << logC;
    const int row_base = o * W.row_o;
    for (int i = tid; i < total; i += nt) {
        const int c = i & (C - 1), e = i >> logC;
        const int row = row_base + e * W.row_e;
        smem[PADI(i)] = (c < cw && row >= W.in_lo && row < W.in_hi) ? in[base + (int64_t)e * elem_stride + c]
                                                                    : make_double2(0.0, 0.0);
    }
    __syncthreads();
    if (do_fwd) tile_fft<1, EPT, CLS>(smem, P, C, logC, -1, tid, nt);
    for (int i = tid; i < total; i += nt) {
        const int c = i & (C - 1), e = i >> logC;
        if (c < cw) {
            const int64_t addr = base + (int64_t)e * elem_stride + c;
            cplx v = smem[PADI(i)];
            if (copy_mode == 1) copy[addr] = v;
            v = colmult_apply(M, v, addr, o, e, (int)c0 + c);
            if (copy_mode == 2) copy[addr] = v;
            smem[PADI(i)] = v;
        }
    }
    __syncthreads();
    tile_fft<1, EPT, CLS>(smem, P, C, logC, +1, tid, nt);
    for (int i = tid; i < total; i += nt) {
        const int c = i & (C - 1), e = i >> logC;
        const int row = row_base + e * W.row_e;
        if (c < cw && row >= W.out_lo && row < W.out_hi) out[base + (int64_t)e * elem_stride + c] = smem[PADI(i)];
    }
}

// The multiplier of one spectrum entry (shared by the tile kernel and the register kernel below).
// o: outer line, e: element of the line, col: global column inside the half spectrum (without col_offset).
__device__ __forceinline__ cplx colmult_apply_table(const ColMult& M, cplx v, double m, int o, int e, int col);
__device__ __forceinline__ cplx colmult_apply(const ColMult& M, cplx v, int64_t addr, int o, int e, int col) {
    if (M.kind == 3) return colmult_apply_table(M, v, M.table[addr], o, e, col);
    const int fi = o + M.fmul * e;                                   // frequency index along x
    const int si = (fi < (M.nx + 1) / 2) ? fi : fi - M.nx;
    const int jy = col + M.col_offset;
    const double kx = TWO_PI * (double)si / ((double)M.nx * M.xh);
    const double ky = TWO_PI * (double)jy / ((double)M.ny * M.yh);
    double m;
    if (M.kind == IPDE_KIND_POISSON) m = (fi == 0 && jy == 0) ? 0.0 : -1.0 / (kx * kx + ky * ky);
    else m = 1.0 / (M.k2c + kx * kx + ky * ky);
    return cscale(v, m);
}

// kind 3 with the table value already in a register
__device__ __forceinline__ cplx colmult_apply_table(const ColMult& M, cplx v, double m, int o, int e, int col) {
    if (M.deriv) {
        // spectral derivative of the far field: multiply by i k (zero on the unpaired Nyquist line)
        double kk;
        if (M.deriv == 1) {
            const int fi = o + M.fmul * e;
            const int si = (fi < (M.nx + 1) / 2) ? fi : fi - M.nx;
            kk = (M.nx % 2 == 0 && fi == M.nx / 2) ? 0.0 : TWO_PI * (double)si / ((double)M.nx * M.xh);
        } else {
            const int jy = col + M.col_offset;
            kk = (jy == M.ny / 2) ? 0.0 : TWO_PI * (double)jy / ((double)M.ny * M.yh);
        }
        v = make_double2(-kk * v.y, kk * v.x);
    }
    return cscale(v, m);
}

// Register-resident version of k_cols_mult for lines of 16, 32 or 64 (fft_short.cuh): forward transform, multiplier
// and inverse transform of a tile with the 8 elements of a thread never leaving its registers except for the two
// exchanges; the next tile's loads are in flight meanwhile.
template <int RA, int RB>
__global__ void __launch_bounds__(128, 4)
k_cols_mult_short(const cplx* __restrict__ tw, const cplx* in, cplx* out, cplx* copy, int64_t elem_stride, int64_t n_inner,
                  int64_t outer_stride, int do_fwd, int copy_mode, ColMult M, RowWindow W, int tiles_per_outer,
                  int ntiles_total) {
    using K = ShortCfg<RA, RB>;
    constexpr int E = K::EPT;
    __shared__ cplx sm[K::NBUF][K::TILE];
    const int tid = threadIdx.x;
    const int c = tid & (K::C - 1), jt = tid >> K::LOGC;
    cplx v[E], nx[K::PREFETCH ? E : 1];
    auto load = [&](int tile, cplx* d) {
        const ShortTile S = short_tile<K::LOGC>(tile, tiles_per_outer, c, n_inner, outer_stride, W.row_o);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int e = do_fwd ? short_elem_a<RA, RB>(jt, s) : short_elem_b<RA, RB>(jt, s);
            const int row = S.row_base + e * W.row_e;
            d[s] = (S.colok && row >= W.in_lo && row < W.in_hi) ? in[S.base + (int64_t)e * elem_stride]
                                                                : make_double2(0.0, 0.0);
        }
    };
    int tile = blockIdx.x;
    if (K::PREFETCH && tile < ntiles_total) load(tile, v);
    for (int par = 0; tile < ntiles_total; tile += gridDim.x, par ^= 1) {
        if (K::PREFETCH) {
            const int next = tile + gridDim.x;
            if (next < ntiles_total) load(next, nx);
        } else {
            load(tile, v);
        }
        const ShortTile S = short_tile<K::LOGC>(tile, tiles_per_outer, c, n_inner, outer_stride, W.row_o);
        // table symbol: issue its loads now, they arrive while the forward transform runs
        double mt[K::PREFETCH ? E : 1];
        const bool table_early = K::PREFETCH && M.kind == 3 && S.colok;
        if (table_early) {
#pragma unroll
            for (int s = 0; s < E; ++s) mt[s] = M.table[S.base + (int64_t)short_elem_b<RA, RB>(jt, s) * elem_stride];
        }
        if (do_fwd) short_fft<RA, RB, -1>(v, sm[0], c, jt, tw);
        if (S.colok) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const int e = short_elem_b<RA, RB>(jt, s);
                const int64_t addr = S.base + (int64_t)e * elem_stride;
                if (copy_mode == 1) copy[addr] = v[s];
                v[s] = table_early ? colmult_apply_table(M, v[s], mt[s], S.o, e, S.col) : colmult_apply(M, v[s], addr, S.o, e, S.col);
                if (copy_mode == 2) copy[addr] = v[s];
            }
        }
        // two buffers: with the forward transform the two exchanges of a tile alternate by themselves, without it
        // consecutive tiles must; one buffer: short_fft synchronises before rewriting it
        short_fft<RB, RA, +1>(v, sm[K::NBUF == 1 ? 0 : (do_fwd ? 1 : par)], c, jt, tw);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int e = short_elem_a<RA, RB>(jt, s);
            const int row = S.row_base + e * W.row_e;
            if (S.colok && row >= W.out_lo && row < W.out_hi) out[S.base + (int64_t)e * elem_stride] = v[s];
        }
        if (K::PREFETCH) {
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = nx[s];
        }
    }
}

template <int RA, int RB>
static int launch_cols_mult_short(FftPlan* p, const cplx* in, cplx* out, cplx* copy, int64_t elem_stride, int64_t n_inner,
                                  int64_t n_outer, int64_t outer_stride, int do_fwd, int copy_mode, const ColMult& M,
                                  const RowWindow& W, cudaStream_t st) {
    using K = ShortCfg<RA, RB>;
    const int64_t tiles = cdiv(n_inner, (int64_t)K::C);
    const int64_t total = tiles * n_outer;
    ARG_CHECK(total < (int64_t)1 << 30, "short column transform: too many tiles");
    KERNEL_TAG("k_cols_mult", (32.0 + (copy_mode ? 16.0 : 0.0) + (M.kind == 3 ? 8.0 : 0.0)) * (double)p->d.n * (double)n_inner *
                                  (double)n_outer);
    const unsigned g = (unsigned)std::min<int64_t>(total, (int64_t)ipde_num_sms() * 4);
    k_cols_mult_short<RA, RB><<<g, 128, 0, st>>>(p->d.tw, in, out, copy, elem_stride, n_inner, outer_stride, do_fwd, copy_mode,
                                                 M, W, (int)tiles, (int)total);
    KERNEL_CHECK();
    return IPDE_OK;
}

template <int EPT, int CLS, int NT>
static int launch_cols_mult_one(FftPlan* p, unsigned g, int t, size_t smem, cudaStream_t st, const cplx* in, cplx* out,
                                cplx* copy, int64_t elem_stride, int64_t n_inner, int64_t outer_stride, int C, int logC,
                                int do_fwd, int copy_mode, const ColMult& M, const RowWindow& W) {
    KERNEL_TAG("k_cols_mult", (32.0 + (copy_mode ? 16.0 : 0.0) + (M.kind == 3 ? 8.0 : 0.0)) * (double)p->d.n * (double)n_inner *
                                  (double)(g / cdiv(n_inner, C)));
    if (smem > 48 * 1024)
        CUDA_TRY(cudaFuncSetAttribute(k_cols_mult<EPT, CLS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cols_mult<EPT, CLS, NT><<<g, t, smem, st>>>(p->d, in, out, copy, elem_stride, n_inner, outer_stride, C, logC, do_fwd,
                                                  copy_mode, M, W);
    KERNEL_CHECK();
    return IPDE_OK;
}

// columns of length p->d.n (must fit a tile with >= 4 adjacent columns and be tile-engine friendly)
static int cols_mult(FftPlan* p, const cplx* in, cplx* out, cplx* copy, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                     int64_t outer_stride, int do_fwd, int copy_mode, const ColMult& M, const RowWindow& W, cudaStream_t st) {
    const int n = p->d.n;
    if (fft_short_enabled() && p->pow2 && fft_short_size(n)) {
#define CALL_SHORT(RA_, RB_) return launch_cols_mult_short<RA_, RB_>(p, in, out, copy, elem_stride, n_inner, n_outer, outer_stride, do_fwd, copy_mode, M, W, st)
        SHORT_DISPATCH(n, CALL_SHORT);
#undef CALL_SHORT
    }
    int C = 1;
    while ((int64_t)C * 2 * n <= IPDE_FFT_SMEM_MAX && C < 16) C *= 2;
    while (C > 1 && C / 2 >= n_inner) C /= 2;
    int logC = 0;
    while ((1 << logC) < C) ++logC;
    const int64_t E = (int64_t)n * C;
    const int ept = E > 4096 ? 16 : 8;
    const size_t smem = fft_tile_smem(E);
    const unsigned g = (unsigned)(cdiv(n_inner, C) * n_outer);
    const int t = fft_tile_threads(n, C, ept);
#define CM(E_, C_, N_) return launch_cols_mult_one<E_, C_, N_>(p, g, t, smem, st, in, out, copy, elem_stride, n_inner, \
                                                              outer_stride, C, logC, do_fwd, copy_mode, M, W)
    if (ept == 8) {
        if (p->pow2) { if (t <= 256) CM(8, 0, 256); else CM(8, 0, 512); }
        else { if (t <= 256) CM(8, 1, 256); else CM(8, 1, 512); }
    } else {
        if (p->pow2) { if (t <= 256) CM(16, 0, 256); else CM(16, 0, 512); }
        else { if (t <= 256) CM(16, 1, 256); else CM(16, 1, 512); }
    }
#undef CM
    return IPDE_OK;
}

static inline bool grid_fused_ok(const GridPlan* g) {
    const FftPlan* inner = (g->n2 == 1) ? g->colA : g->colB;
    return inner && inner->fast && (int64_t)inner->d.n * 4 <= IPDE_FFT_SMEM_MAX;
}

// spec_in (row transforms done, columns not) -> [column FFT, symbol, inverse column FFT] -> spec_out, ready for
// the inverse row transforms.  spec_copy (may be null) receives the 2-D spectrum, before (copy_mode 1) or after
// (2) the multiplication.  spec_in is overwritten when the column transform is split.
int grid_cols_symbol(GridPlan* g, cplx* spec_in, cplx* spec_out, cplx* spec_copy, int copy_mode, const ColMult& M0,
                     cudaStream_t st, const int* win, const SlabScatter* sc) {
    const int64_t P = g->pitch;
    // win = {in_lo, in_hi, out_lo, out_hi}: rows of spec_in outside [in_lo, in_hi) are zero (not read), rows of
    // spec_out outside [out_lo, out_hi) are not needed (not written)
    const int in_lo = win ? win[0] : 0, in_hi = win ? win[1] : g->nx;
    const int out_lo = win ? win[2] : 0, out_hi = win ? win[3] : g->nx;
    ColMult M = M0;
    M.nx = g->nx; M.xh = g->xh;
    if (M.ny <= 0) { M.ny = g->ny; M.yh = g->yh; M.col_offset = 0; }
    if (!spec_copy) copy_mode = 0;
    if (g->n2 == 1) {
        if (sc && sc->R) {
            ipde_set_error("grid_cols_symbol: scattered stores need a four-step column split (nx = %d)", g->nx);
            return IPDE_ERR_UNSUPPORTED;
        }
        M.fmul = 1;
        RowWindow W = {0, 1, in_lo, in_hi, out_lo, out_hi};
        return cols_mult(g->colA, spec_in, spec_out, spec_copy, P, P, 1, 0, 1, copy_mode, M, W, st);
    }
    M.fmul = g->n1;
    // four-step: element e of line b of the first pass is row b + n2*e
    RowWindow Wa = {1, g->n2, in_lo, in_hi, 0, g->nx};
    int rc = fft_cols_windowed(g->colA, spec_in, spec_in, (int64_t)g->n2 * P, P, g->n2, P, -1, 1.0, 1, g->tw_big, Wa, st);
    if (rc) return rc;
    rc = cols_mult(g->colB, spec_in, spec_out, spec_copy, P, P, g->n1, (int64_t)g->n2 * P, 1, copy_mode, M, row_window_all(), st);
    if (rc) return rc;
    RowWindow Wb = {1, g->n2, 0, g->nx, out_lo, out_hi};
    return fft_cols_windowed(g->colA, spec_out, spec_out, (int64_t)g->n2 * P, P, g->n2, P, +1, 1.0, 2, g->tw_big, Wb, st, sc);
}

// 2-D spectrum (as left by grid_cols_symbol's copy) -> symbol -> inverse column FFT -> spec_out
int grid_cols_symbol_inverse(GridPlan* g, const cplx* spec2d, cplx* spec_out, const ColMult& M0, cudaStream_t st,
                             const int* win) {
    const int64_t P = g->pitch;
    const int out_lo = win ? win[0] : 0, out_hi = win ? win[1] : g->nx;
    ColMult M = M0;
    M.nx = g->nx; M.xh = g->xh;
    if (M.ny <= 0) { M.ny = g->ny; M.yh = g->yh; M.col_offset = 0; }
    if (g->n2 == 1) {
        M.fmul = 1;
        RowWindow W = {0, 1, 0, g->nx, out_lo, out_hi};
        return cols_mult(g->colA, spec2d, spec_out, nullptr, P, P, 1, 0, 0, 0, M, W, st);
    }
    M.fmul = g->n1;
    int rc = cols_mult(g->colB, spec2d, spec_out, nullptr, P, P, g->n1, (int64_t)g->n2 * P, 0, 0, M, row_window_all(), st);
    if (rc) return rc;
    RowWindow Wb = {1, g->n2, 0, g->nx, out_lo, out_hi};
    return fft_cols_windowed(g->colA, spec_out, spec_out, (int64_t)g->n2 * P, P, g->n2, P, +1, 1.0, 2, g->tw_big, Wb, st);
}

// the whole scalar grid solve: fc (real) -> uch (2-D spectrum of the solution) and uc (real)
int grid_solve_scalar_fused(GridPlan* g, int kind, double k, const double* fc, cplx* uch, cplx* spec_work, double* uc,
                            cudaStream_t st) {
    int rc = fft_r2c_rows(g->row, fc, uch, g->nx, g->ny, g->pitch, st);
    if (rc) return rc;
    if (grid_fused_ok(g)) {
        ColMult M;
        M.kind = kind; M.k2c = k * k; M.table = nullptr; M.fmul = 1; M.ny = 0; M.col_offset = 0; M.deriv = 0;
        rc = grid_cols_symbol(g, uch, spec_work, uch, 2, M, st);
        if (rc) return rc;
    } else {
        rc = grid_cols_forward(g, uch, uch, st);
        if (!rc) rc = grid_scalar_multiply(g, uch, kind, k, st);
        if (!rc) rc = grid_cols_inverse(g, uch, spec_work, st);
        if (rc) return rc;
    }
    return fft_c2r_rows(g->row, spec_work, uc, g->nx, g->pitch, g->ny, 1.0 / ((double)g->nx * (double)g->ny) * 2.0, st);
}

// ------------------------------------------------------------------------------------------
// compact physical vector -> zero-filled grid, times the cutoff
__global__ void k_scatter_cutoff(double* __restrict__ fc, const double* __restrict__ f, const int64_t* __restrict__ lin,
                                 const double* __restrict__ step, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        fc[lin[i]] = f[i] * step[i];
}

static inline int grid_blocks(int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, threads), (int64_t)ipde_num_sms() * 8));
}

int grid_scatter_cutoff(GridPlan* g, const double* f_compact, const int64_t* phys_lin, const double* step_compact,
                        int64_t n_phys, double* fc, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(fc, 0, sizeof(double) * (size_t)g->nx * g->ny, st));
    if (n_phys > 0) {
        k_scatter_cutoff<<<grid_blocks(n_phys, 256), 256, 0, st>>>(fc, f_compact, phys_lin, step_compact, n_phys);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}

// sum reduction (deterministic two-stage) and demean
__global__ void k_partial_sum(const double* __restrict__ x, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += x[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
__global__ void k_final_sum(double* __restrict__ partial, int nparts) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[nparts] = v;
    }
}
__global__ void k_demean(double* __restrict__ f, const double* __restrict__ bumpy, int64_t n, const double* __restrict__ total,
                         double weight) {
    const double m = total[0] * weight;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        f[i] -= m * bumpy[i];
}

int grid_demean(GridPlan* g, double* f, const double* bumpy, cudaStream_t st) {
    const int64_t n = (int64_t)g->nx * g->ny;
    const int parts = ipde_num_sms() * 4;
    if (!g->reduce_buf) CUDA_TRY(cudaMalloc(&g->reduce_buf, sizeof(double) * (parts + 1)));
    k_partial_sum<<<parts, 256, 0, st>>>(f, n, g->reduce_buf);
    KERNEL_CHECK();
    k_final_sum<<<1, 256, 0, st>>>(g->reduce_buf, parts);
    KERNEL_CHECK();
    k_demean<<<grid_blocks(n, 256), 256, 0, st>>>(f, bumpy, n, g->reduce_buf + parts, g->xh * g->yh);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// spectral multipliers on the half spectrum; rows may be in four-step order
__device__ __forceinline__ int freq_of_row(int pos, int n1, int n2) {
    if (n2 == 1) return pos;
    const int k1 = pos / n2, k2 = pos % n2;
    return k1 + n1 * k2;
}
__device__ __forceinline__ double wavenumber(int i, int n, double h) {
    const int s = (i < (n + 1) / 2) ? i : i - n;     // numpy.fft.fftfreq ordering
    return TWO_PI * (double)s / ((double)n * h);
}

__global__ void k_scalar_multiply(cplx* __restrict__ spec, int nx, int ny, int64_t pitch, int n1, int n2, double xh,
                                  double yh, int kind, double k2const) {
    const int64_t total = (int64_t)nx * pitch;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int pos = (int)(t / pitch), j = (int)(t % pitch);
        const int i = freq_of_row(pos, n1, n2);
        const double kx = wavenumber(i, nx, xh), ky = wavenumber(j, ny, yh);
        double m;
        if (kind == IPDE_KIND_POISSON) {
            const double lap = -(kx * kx + ky * ky);
            m = (i == 0 && j == 0) ? 0.0 : 1.0 / lap;
        } else if (kind == IPDE_KIND_LAPLACIAN) {
            m = -(kx * kx + ky * ky);          // ebdyc.laplacian: fch * self.lap (ipde/ebdy_collection.py:771-774)
        } else {
            m = 1.0 / (k2const + kx * kx + ky * ky);
        }
        spec[t] = cscale(spec[t], m);
    }
}

int grid_scalar_multiply(GridPlan* g, cplx* spec, int kind, double k, cudaStream_t st) {
    const int64_t total = (int64_t)g->nx * g->pitch;
    k_scalar_multiply<<<grid_blocks(total, 256), 256, 0, st>>>(spec, g->nx, g->ny, g->pitch, g->n1, g->n2, g->xh, g->yh,
                                                               kind, k * k);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// 4th-order centred differences, zero in the 2-cell border (reference derivatives.py:3-23)
__global__ void k_fd4(const double* __restrict__ f, double* __restrict__ fx, double* __restrict__ fy, int nx, int ny,
                      double iahx, double iahy) {
    const int64_t total = (int64_t)nx * ny;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / ny), j = (int)(t % ny);
        double dx = 0.0, dy = 0.0;
        if (i >= 2 && i < nx - 2)
            dx = -(f[t + 2 * (int64_t)ny] - 8.0 * f[t + ny] + 8.0 * f[t - ny] - f[t - 2 * (int64_t)ny]) * iahx;
        if (j >= 2 && j < ny - 2) dy = -(f[t + 2] - 8.0 * f[t + 1] + 8.0 * f[t - 1] - f[t - 2]) * iahy;
        if (fx) fx[t] = dx;
        if (fy) fy[t] = dy;
    }
}

int grid_fd4(GridPlan* g, const double* f, double* fx, double* fy, cudaStream_t st) {
    const int64_t total = (int64_t)g->nx * g->ny;
    k_fd4<<<grid_blocks(total, 256), 256, 0, st>>>(f, fx, fy, g->nx, g->ny, 1.0 / (12.0 * g->xh), 1.0 / (12.0 * g->yh));
    KERNEL_CHECK();
    return IPDE_OK;
}

// periodic tensor-product cubic Lagrange interpolation (fast_interp.interp2d(k=3, p=[True,True]))
__global__ void k_interp_cubic(const double* __restrict__ f, int nx, int ny, double x0, double y0, double xh, double yh,
                               const double* __restrict__ px, const double* __restrict__ py, int64_t npts,
                               double* __restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npts; p += (int64_t)gridDim.x * blockDim.x) {
        const double sx = (px[p] - x0) / xh, sy = (py[p] - y0) / yh;
        const int ix = (int)floor(sx) - 1, iy = (int)floor(sy) - 1;
        const double ux = sx - ix, uy = sy - iy;
        double wx[4], wy[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double wa = 1.0, wb = 1.0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (a != b) {
                    wa *= (ux - b) / (double)(a - b);
                    wb *= (uy - b) / (double)(a - b);
                }
            wx[a] = wa; wy[a] = wb;
        }
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ii = ((ix + a) % nx + nx) % nx;
            double row = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int jj = ((iy + b) % ny + ny) % ny;
                row += f[(int64_t)ii * ny + jj] * wy[b];
            }
            acc += row * wx[a];
        }
        out[p] = acc;
    }
}

int grid_interp_cubic(GridPlan* g, const double* f, double x0, double y0, const double* px, const double* py,
                      int64_t npts, double* out, cudaStream_t st) {
    if (npts <= 0) return IPDE_OK;
    k_interp_cubic<<<grid_blocks(npts, 128), 128, 0, st>>>(f, g->nx, g->ny, x0, y0, g->xh, g->yh, px, py, npts, out);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// finalize: compact[g] = uc[phys_lin[g]] ; compact[pna_pos[i]] += gu[i] ; compact[ia_pos[i]] = radial value
__global__ void k_gather(const double* __restrict__ uc, const int64_t* __restrict__ lin, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = uc[lin[i]];
}
__global__ void k_index_add(double* __restrict__ out, const int64_t* __restrict__ pos, const double* __restrict__ v, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[pos[i]] += v[i];
}
__global__ void k_index_set(double* __restrict__ out, const int64_t* __restrict__ pos, const double* __restrict__ v, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[pos[i]] = v[i];
}

int grid_gather(const double* uc, const int64_t* lin, int64_t n, double* out, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_gather<<<grid_blocks(n, 256), 256, 0, st>>>(uc, lin, n, out);
    KERNEL_CHECK();
    return IPDE_OK;
}
int grid_index_add(double* out, const int64_t* pos, const double* v, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_index_add<<<grid_blocks(n, 256), 256, 0, st>>>(out, pos, v, n);
    KERNEL_CHECK();
    return IPDE_OK;
}
int grid_index_set(double* out, const int64_t* pos, const double* v, int64_t n, cudaStream_t st) {
    if (n <= 0) return IPDE_OK;
    k_index_set<<<grid_blocks(n, 256), 256, 0, st>>>(out, pos, v, n);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Stokes multiplier (mu = 1), reference ipde/solvers/multi_boundary/stokes.py:34-50:
//   ph = ilap (ikx fuh + iky fvh) ; uh = ilap (ikx ph - fuh) ; vh = ilap (iky ph - fvh)
// followed there by ifft2(.).real.  Taking the real part symmetrises the spectrum; away from the Nyquist
// row/column that changes nothing, on them it averages the expression over the two signs of the Nyquist
// wavenumber ("Nyquist modes").
struct Uvp { cplx u, v, p; };
__device__ __forceinline__ Uvp stokes_symbol(double kx, double ky, cplx fu, cplx fv, bool zero_mode) {
    Uvp r;
    const double lap = -(kx * kx + ky * ky);
    const double il = zero_mode ? 0.0 : 1.0 / lap;
    // ikx*fu + iky*fv
    cplx d = make_double2(-(kx * fu.y + ky * fv.y), kx * fu.x + ky * fv.x);
    r.p = cscale(d, il);
    cplx ikxp = make_double2(-kx * r.p.y, kx * r.p.x);
    cplx ikyp = make_double2(-ky * r.p.y, ky * r.p.x);
    r.u = cscale(csub(ikxp, fu), il);
    r.v = cscale(csub(ikyp, fv), il);
    return r;
}

__global__ void k_stokes_multiply(cplx* __restrict__ uh, cplx* __restrict__ vh, cplx* __restrict__ ph, int nx, int ny,
                                  int64_t pitch, int n1, int n2, double xh, double yh) {
    const int64_t total = (int64_t)nx * pitch;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int pos = (int)(t / pitch), j = (int)(t % pitch);
        const int i = freq_of_row(pos, n1, n2);
        const double kx = wavenumber(i, nx, xh), ky = wavenumber(j, ny, yh);
        const bool xn = (nx % 2 == 0 && i == nx / 2), yn = (j == ny / 2);
        const cplx fu = uh[t], fv = vh[t];
        Uvp a = stokes_symbol(kx, ky, fu, fv, i == 0 && j == 0);
        if (xn || yn) {
            Uvp b = stokes_symbol(xn ? -kx : kx, yn ? -ky : ky, fu, fv, false);
            a.u = cscale(cadd(a.u, b.u), 0.5);
            a.v = cscale(cadd(a.v, b.v), 0.5);
            a.p = cscale(cadd(a.p, b.p), 0.5);
        }
        uh[t] = a.u; vh[t] = a.v; ph[t] = a.p;
    }
}

int grid_stokes_multiply(GridPlan* g, cplx* uh, cplx* vh, cplx* ph, cudaStream_t st) {
    const int64_t total = (int64_t)g->nx * g->pitch;
    k_stokes_multiply<<<grid_blocks(total, 256), 256, 0, st>>>(uh, vh, ph, g->nx, g->ny, g->pitch, g->n1, g->n2, g->xh, g->yh);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Spectra of the grid stress of the Stokes solve, reference ipde/solvers/multi_boundary/vector.py:68-73 with the
// spectral derivative of ipde/derivatives.py:25-28: d/dx f = ifft2(i kx fft2(f)).real.  Taking the real part of
// that product removes the (self-conjugate) Nyquist row for d/dx and the Nyquist column for d/dy exactly, so the
// derivative symbols are zero there.
//   txx = 2 u_x - p ;  txy = u_y + v_x ;  tyy = 2 v_y - p
__global__ void k_stokes_stress(const cplx* __restrict__ uh, const cplx* __restrict__ vh, const cplx* __restrict__ ph,
                                cplx* __restrict__ txx, cplx* __restrict__ txy, cplx* __restrict__ tyy, int nx, int ny,
                                int64_t pitch, int n1, int n2, double xh, double yh) {
    const int64_t total = (int64_t)nx * pitch;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int pos = (int)(t / pitch), j = (int)(t % pitch);
        const int i = freq_of_row(pos, n1, n2);
        const double kx = (nx % 2 == 0 && i == nx / 2) ? 0.0 : wavenumber(i, nx, xh);
        const double ky = (j == ny / 2) ? 0.0 : wavenumber(j, ny, yh);
        const cplx u = uh[t], v = vh[t], p = ph[t];
        // i k z = (-k z.y, k z.x)
        txx[t] = make_double2(-2.0 * kx * u.y - p.x, 2.0 * kx * u.x - p.y);
        txy[t] = make_double2(-ky * u.y - kx * v.y, ky * u.x + kx * v.x);
        tyy[t] = make_double2(-2.0 * ky * v.y - p.x, 2.0 * ky * v.x - p.y);
    }
}

int grid_stokes_stress(GridPlan* g, const cplx* uh, const cplx* vh, const cplx* ph, cplx* txx, cplx* txy, cplx* tyy,
                       cudaStream_t st) {
    const int64_t total = (int64_t)g->nx * g->pitch;
    k_stokes_stress<<<grid_blocks(total, 256), 256, 0, st>>>(uh, vh, ph, txx, txy, tyy, g->nx, g->ny, g->pitch, g->n1, g->n2,
                                                             g->xh, g->yh);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// slab decomposition helpers
// Column block q of the half spectrum is [q*Pc, (q+1)*Pc); a chunk is the sub-block [koff, koff + Pk) of every block
// (koff = 0, Pk = Pc: the whole block).  Chunks let the all-to-all of one part overlap the column pass of another.
__global__ void k_slab_pack(const cplx* __restrict__ in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk,
                            cplx* __restrict__ out) {
    const int64_t total = (int64_t)R * rows * Pk;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kc = t % Pk, l = (t / Pk) % rows, q = t / (Pk * rows);
        const int64_t k = q * Pc + koff + kc;
        out[t] = (k < P) ? in[l * P + k] : make_double2(0.0, 0.0);
    }
}
__global__ void k_slab_unpack(const cplx* __restrict__ in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk,
                              cplx* __restrict__ out) {
    const int64_t total = (int64_t)R * rows * Pk;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kc = t % Pk, l = (t / Pk) % rows, q = t / (Pk * rows);
        const int64_t k = q * Pc + koff + kc;
        if (k < P) out[l * P + k] = in[t];
    }
}
int grid_slab_pack(const cplx* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, cplx* out,
                   cudaStream_t st) {
    ARG_CHECK(koff >= 0 && Pk > 0 && koff + Pk <= Pc, "slab pack: chunk outside the column block");
    const int64_t total = (int64_t)R * rows * Pk;
    KERNEL_TAG("k_slab_pack", 32.0 * (double)total);
    k_slab_pack<<<grid_blocks(total, 256), 256, 0, st>>>(in, rows, P, R, Pc, koff, Pk, out);
    KERNEL_CHECK();
    return IPDE_OK;
}
int grid_slab_unpack(const cplx* in, int64_t rows, int64_t P, int R, int64_t Pc, int64_t koff, int64_t Pk, cplx* out,
                     cudaStream_t st) {
    ARG_CHECK(koff >= 0 && Pk > 0 && koff + Pk <= Pc, "slab unpack: chunk outside the column block");
    const int64_t total = (int64_t)R * rows * Pk;
    KERNEL_TAG("k_slab_unpack", 32.0 * (double)total);
    k_slab_unpack<<<grid_blocks(total, 256), 256, 0, st>>>(in, rows, P, R, Pc, koff, Pk, out);
    KERNEL_CHECK();
    return IPDE_OK;
}
int grid_cols_symbol_slab_scatter(GridPlan* g, cplx* spec_in, cplx* spec_work, int kind, double k, int ny_global, double yh,
                                  int col_offset, const SlabScatter& sc, cudaStream_t st) {
    ARG_CHECK(grid_fused_ok(g), "column length of the slab plan is not supported by the fused column pass");
    ARG_CHECK(sc.R >= 1 && sc.R <= IPDE_SLAB_MAX_RANKS && sc.rows * sc.R == g->nx && sc.Pc == g->pitch,
              "slab scatter: rows * ranks must equal nx and Pc the pitch of the column plan");
    ColMult M;
    M.kind = kind; M.k2c = k * k; M.table = nullptr; M.fmul = 1; M.deriv = 0;
    M.ny = ny_global; M.yh = yh; M.col_offset = col_offset;
    return grid_cols_symbol(g, spec_in, spec_work, nullptr, 0, M, st, nullptr, &sc);
}
int grid_cols_symbol_slab(GridPlan* g, cplx* spec_in, cplx* spec_out, cplx* spec_copy, int copy_mode, int kind, double k,
                          int ny_global, double yh, int col_offset, cudaStream_t st) {
    ARG_CHECK(grid_fused_ok(g), "column length of the slab plan is not supported by the fused column pass");
    ColMult M;
    M.kind = kind; M.k2c = k * k; M.table = nullptr; M.fmul = 1; M.deriv = 0;
    M.ny = ny_global; M.yh = yh; M.col_offset = col_offset;
    return grid_cols_symbol(g, spec_in, spec_out, spec_copy, copy_mode, M, st);
}
