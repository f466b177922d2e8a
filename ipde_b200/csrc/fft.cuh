// Hand-written FP64 Stockham FFT for sm_100a: plans + the shared-memory tile engine.
// No cuFFT anywhere.  See fft.cu for the kernels and DESIGN.md for the pass structure.
//
// Tile engine.  A CTA holds a tile of C lines of length n (n*C <= IPDE_FFT_SMEM_MAX complex numbers) in ONE
// shared-memory buffer.  A radix-R pass is done in place through registers: every thread reads the R inputs of
// its butterflies, the CTA synchronises, every thread twiddles, transforms and writes its R outputs to the
// autosort (Stockham) positions.  Each thread owns 16 elements per pass (16/R butterflies), so a tile of 4096
// elements runs on 256 threads and two CTAs fit an SM: one CTA's global loads overlap the other's passes.
// Shared memory is skewed by one 16-byte slot every 8 (PADI) so that the stride-R writes of the first passes
// are bank-conflict free.  Twiddles: one table load per butterfly, the remaining powers by a depth-log2(R)
// product tree (<= 3 roundings for R = 8).
#pragma once
#include "common.cuh"

#define IPDE_FFT_MAX_FACTORS 24
#define IPDE_FFT_MAX_RADIX 13
// capacity of one tile in complex elements: 8192 * 9/8 * 16 B = 144 KB of shared memory, 512 threads
#define IPDE_FFT_SMEM_MAX 8192
#define PADI(u) ((u) + ((u) >> 3))

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
    bool fast = true;             // all radices in {2,3,4,5,8}: tile engine; otherwise fft_generic.cu
    bool pow2 = false;            // radices 4 and 8 only (lean kernel instantiation)
    bool any = false;             // has a prime factor > 13: global-memory passes with an O(R^2) butterfly
    size_t work_elems = 0;
    FftPlan* half = nullptr;      // plan of length m = n/R for the split long-line path (m a register-kernel size)
    int splitR = 0;               // R of that split (0: not split)
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
// follows `sign`), tw_mode 2 multiplies the *input* likewise.  n_in <= n: only the first n_in elements of
// every input line are read (the rest are zero); only the first n_out output elements are written.
int fft_cols(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
             int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
             cudaStream_t st);
// Row windows for zero-padded data (free-space evaluator): element e of outer line o sits in row
// o*row_o + e*row_e of the array; rows outside [in_lo, in_hi) are known to be zero and are not read, rows outside
// [out_lo, out_hi) are not needed and are not written.
struct RowWindow {
    int row_o, row_e;
    int in_lo, in_hi, out_lo, out_hi;
};
static inline RowWindow row_window_all() { RowWindow w = {0, 0, 0, 1, 0, 1}; return w; }

// Slab-decomposed grids (ipde_b200/distributed.py, SlabGridSolverP2P): the transposes between row slabs and column
// slabs are done by the STORES of the transform kernels, straight into the peers' buffers over NVLink (CUDA IPC
// mappings), instead of pack + all-to-all + unpack.
//   forward   rank p owns rows [row0, row0 + rows_local): entry (local row l, column k) of the row spectra goes to
//             the column slab (nx, Pc) of rank q = k / Pc, position [(row0 + l) * Pc + k - q Pc]
//   backward  rank p owns columns [col0, col0 + Pc): entry (global row i, local column c) of the column pass goes to
//             the row slab (rows, P) of rank q = i / rows, position [(i - q rows) * P + col0 + c]   (col0 + c < P)
#define IPDE_SLAB_MAX_RANKS 8
struct SlabScatter {
    int R;                 // 0: plain store
    int rows;              // rows per rank
    int64_t Pc, P;         // columns per rank, valid columns (ny/2 + 1)
    int64_t row0, col0;    // first global row (forward) / first global column (backward) of this rank
    cplx* base[IPDE_SLAB_MAX_RANKS];
};
static inline SlabScatter slab_scatter_none() { SlabScatter s = {}; return s; }
__device__ __forceinline__ cplx* slab_fwd_addr(const SlabScatter& S, int64_t l, int k) {
    const int q = (int)(k / S.Pc);
    return S.base[q] + (S.row0 + l) * S.Pc + (k - q * S.Pc);
}
__device__ __forceinline__ cplx* slab_bwd_addr(const SlabScatter& S, int row, int c) {
    const int q = row / S.rows;
    return S.base[q] + (int64_t)(row - q * S.rows) * S.P + S.col0 + c;
}

int fft_cols_windowed(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                      int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, RowWindow W,
                      cudaStream_t st, const SlabScatter* sc = nullptr);

// Real <-> half-complex rows (length nr = 2*p->d.n reals <-> nr/2+1 complex), plan is for nr/2.
int fft_r2c_rows(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 cudaStream_t st);
int fft_c2r_rows(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 double scale, cudaStream_t st);
// the same when only the first nz entries of every input row can be nonzero (zero-padded spectra)
// (and only the first out_nz reals of every output row are needed: 0 = all)
// mask (optional): bit b of word mask[row * mask_pitch + (b >> 5)] set <=> the reals [64 b, 64 b + 64) of that output
// row are needed; the other blocks may be left unwritten (honoured by the long-row and register kernels)
int fft_c2r_rows_nz(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                    double scale, int nz, cudaStream_t st, int64_t out_nz = 0, const uint32_t* mask = nullptr,
                    int64_t mask_pitch = 0);
// r2c of the local rows of a slab with the forward transpose on the stores (SlabScatter); `scratch` (batch, scratch_pitch)
// holds the sub-transforms of long rows.  IPDE_ERR_UNSUPPORTED for row lengths without a register kernel.
int fft_r2c_rows_scatter(FftPlan* p, const double* in, cplx* scratch, int64_t batch, int64_t in_pitch, int64_t scratch_pitch,
                         const SlabScatter& sc, cudaStream_t st);
// real rows whose entries beyond in_nz are zero (zero-padded rows)
int fft_r2c_rows_nz(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                    int64_t in_nz, cudaStream_t st);

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

// 16 = 4 x 4 (input index 4 n1 + n2, output index k1 + 4 k2), constant inner twiddles w16^(n2 k1)
template <>
__device__ __forceinline__ void dft_small<16>(cplx* v, int sign, const FftPlanDev& P) {
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    const double sg = (sign < 0) ? -1.0 : 1.0;
    cplx y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        cplx a[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
        dft_small<4>(a, sign, P);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = a[k1];
    }
    const cplx w1 = make_double2(c1, sg * s1), w2 = make_double2(h, sg * h), w3 = make_double2(s1, sg * c1);
    const cplx w6 = make_double2(-h, sg * h), w9 = make_double2(-c1, -sg * s1);
    y[1][1] = cmul(y[1][1], w1); y[1][2] = cmul(y[1][2], w2); y[1][3] = cmul(y[1][3], w3);
    y[2][1] = cmul(y[2][1], w2); y[2][2] = (sign < 0) ? cmulni(y[2][2]) : cmuli(y[2][2]); y[2][3] = cmul(y[2][3], w6);
    y[3][1] = cmul(y[3][1], w3); y[3][2] = cmul(y[3][2], w6); y[3][3] = cmul(y[3][3], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        cplx b[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
        dft_small<4>(b, sign, P);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = b[k2];
    }
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
template <>
__device__ __forceinline__ void dft_small<3>(cplx* v, int sign, const FftPlanDev& P) {
    const double s3 = 0.86602540378443864676;     // sin(2 pi / 3)
    const cplx a = cadd(v[1], v[2]);
    const cplx d = cscale(csub(v[1], v[2]), (sign < 0) ? -s3 : s3);
    const cplx r = make_double2(v[0].x - 0.5 * a.x, v[0].y - 0.5 * a.y);
    v[0] = cadd(v[0], a);
    v[1] = make_double2(r.x - d.y, r.y + d.x);    // r + i d
    v[2] = make_double2(r.x + d.y, r.y - d.x);    // r - i d
}

template <>
__device__ __forceinline__ void dft_small<5>(cplx* v, int sign, const FftPlanDev& P) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2 pi/5), cos(4 pi/5)
    double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;          // sin(2 pi/5), sin(4 pi/5)
    if (sign < 0) { s1 = -s1; s2 = -s2; }
    const cplx a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const cplx b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const cplx r1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const cplx r2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const cplx i1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const cplx i2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = make_double2(r1.x - i1.y, r1.y + i1.x);
    v[4] = make_double2(r1.x + i1.y, r1.y - i1.x);
    v[2] = make_double2(r2.x - i2.y, r2.y + i2.x);
    v[3] = make_double2(r2.x + i2.y, r2.y - i2.x);
}
template <> __device__ __forceinline__ void dft_small<7>(cplx* v, int s, const FftPlanDev& P) { dft_generic<7>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<11>(cplx* v, int s, const FftPlanDev& P) { dft_generic<11>(v, s, P); }
template <> __device__ __forceinline__ void dft_small<13>(cplx* v, int s, const FftPlanDev& P) { dft_generic<13>(v, s, P); }

// Tile layouts: element e of line c lives at logical slot
//   LAYOUT 0 (rows):    c*n + e      (lines contiguous, as in global memory)
//   LAYOUT 1 (columns): e*C + c      (C adjacent lines interleaved, C a power of two)
// and at physical slot PADI(logical).
template <int LAYOUT>
__device__ __forceinline__ int tile_slot(int e, int c, int n, int logC) {
    return PADI(LAYOUT == 0 ? c * n + e : (e << logC) + c);
}

// One in-place radix-R Stockham pass over the tile.  Begins and ends with the tile consistent; contains
// two __syncthreads().  Every thread owns EPT elements per pass (IPT = ceil(EPT/R) butterflies).
//   LAYOUT 1: butterfly index it = tid + q*nt -> (c = it % C, j = it / C), C a power of two;
//   LAYOUT 0: the CTA is split into C groups of G = ceil(n/EPT) threads, group c works on line c
//             (c = tid / G is computed once per kernel: no integer division per butterfly).
struct TileThread {
    int c;      // LAYOUT 0: my line (or -1: idle thread)
    int lt;     // LAYOUT 0: my index inside the group
    int G;      // LAYOUT 0: threads per line
};

template <int R, int LAYOUT, int EPT>
__device__ __forceinline__ void tile_pass(cplx* __restrict__ s, const FftPlanDev& P, int Ns, int C, int logC, int sign,
                                          int tid, int nt, const TileThread& T) {
    constexpr int IPT = (EPT + R - 1) / R;
    const int n = P.n;
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const bool ns_pow2 = (Ns & (Ns - 1)) == 0;
    const int items = (LAYOUT == 1) ? (m << logC) : m;
    const int stride = (LAYOUT == 1) ? nt : T.G;
    const int first = (LAYOUT == 1) ? tid : T.lt;
    const bool live = (LAYOUT == 1) || (T.c >= 0);
    // PADI(x + d) == PADI(x) + d*9/8 whenever d is a multiple of 8, and == PADI(x) + d when x is a multiple of 8
    // and d < 8: the R slots of a butterfly are then equally spaced in the skewed buffer
    const int ls = (LAYOUT == 1) ? (m << logC) : m;
    const int ss = (LAYOUT == 1) ? (Ns << logC) : Ns;
    const bool lin_l = (ls & 7) == 0;
    const int step_l = ls + (ls >> 3);
    const bool first8 = (LAYOUT == 0) && (Ns == 1) && (R == 8) && ((n & 7) == 0);
    const bool lin_s = ((ss & 7) == 0) || first8;
    const int step_s = first8 ? 1 : ss + (ss >> 3);
    cplx v[IPT][R];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int it = first + q * stride;
        if (live && it < items) {
            const int c = (LAYOUT == 1) ? (it & (C - 1)) : T.c;
            const int j = (LAYOUT == 1) ? (it >> logC) : it;
            if (lin_l) {
                const int a = tile_slot<LAYOUT>(j, c, n, logC);
#pragma unroll
                for (int t = 0; t < R; ++t) v[q][t] = s[a + t * step_l];
            } else {
#pragma unroll
                for (int t = 0; t < R; ++t) v[q][t] = s[tile_slot<LAYOUT>(j + t * m, c, n, logC)];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int it = first + q * stride;
        if (live && it < items) {
            const int c = (LAYOUT == 1) ? (it & (C - 1)) : T.c;
            const int j = (LAYOUT == 1) ? (it >> logC) : it;
            const int k = ns_pow2 ? (j & (Ns - 1)) : (j % Ns);
            if (Ns > 1) {
                cplx w[R];
                w[1] = P.tw[k * tstep];
                if (sign > 0) w[1] = cconj(w[1]);
#pragma unroll
                for (int t = 2; t < R; ++t) w[t] = cmul(w[t >> 1], w[t - (t >> 1)]);
#pragma unroll
                for (int t = 1; t < R; ++t) v[q][t] = cmul(v[q][t], w[t]);
            }
            dft_small<R>(v[q], sign, P);
            const int j0 = (j - k) * R + k;
            if (lin_s) {
                const int a = tile_slot<LAYOUT>(j0, c, n, logC);
#pragma unroll
                for (int t = 0; t < R; ++t) s[a + t * step_s] = v[q][t];
            } else {
#pragma unroll
                for (int t = 0; t < R; ++t) s[tile_slot<LAYOUT>(j0 + t * Ns, c, n, logC)] = v[q][t];
            }
        }
    }
    __syncthreads();
}

template <int LAYOUT, int EPT>
__device__ __forceinline__ TileThread tile_thread(int n, int C, int tid) {
    TileThread T;
    T.G = (n + EPT - 1) / EPT;
    if (LAYOUT == 0) {
        const int c = tid / T.G;
        T.lt = tid - c * T.G;
        T.c = (c < C) ? c : -1;
    } else {
        T.c = 0; T.lt = tid;
    }
    return T;
}

// Full transform of the tile, in place.  Caller must have synchronised after filling the tile.
// CLS 0: n is a power of two >= 4 (radices 4 and 8 only -- the lean instantiation every power-of-two grid uses);
// CLS 1: radices 2, 3, 4, 5, 8.
template <int LAYOUT, int EPT, int CLS>
__device__ __forceinline__ void tile_fft(cplx* s, const FftPlanDev& P, int C, int logC, int sign, int tid, int nt) {
    const TileThread T = tile_thread<LAYOUT, EPT>(P.n, C, tid);
    int Ns = 1;
    for (int f = 0; f < P.nfac; ++f) {
        const int R = P.fac[f];
        if (CLS == 0) {
            if (R == 8) tile_pass<8, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T);
            else tile_pass<4, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T);
        } else {
            switch (R) {
                case 2: tile_pass<2, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T); break;
                case 3: tile_pass<3, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T); break;
                case 4: tile_pass<4, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T); break;
                case 5: tile_pass<5, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T); break;
                case 8: tile_pass<8, LAYOUT, EPT>(s, P, Ns, C, logC, sign, tid, nt, T); break;
                default: break;
            }
        }
        Ns *= R;
    }
}

// threads for a tile of `elems` complex numbers (multiple of 32, >= 32)
static inline int fft_tile_threads(int n, int C, int ept) {
    int64_t t = (int64_t)((n + ept - 1) / ept) * C;
    t = (t + 31) / 32 * 32;
    if (t < 32) t = 32;
    return (int)t;
}
static inline size_t fft_tile_smem(int64_t elems) { return sizeof(cplx) * (size_t)(PADI(elems) + 1); }

// ---- cp.async helpers (16-byte global -> shared copies, zero fill when src_bytes == 0) for the double-buffered
// persistent tile kernels
__device__ __forceinline__ void cp_async16(cplx* smem_dst, const cplx* gmem_src, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
bool fft_pipeline_enabled();
