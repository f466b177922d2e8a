// FP64 Stockham FFT kernels (shared-memory tile engine of fft.cuh, mixed radix 2/3/4/5/7/8/11/13) and a
// global-memory pass fallback for lines that do not fit one CTA's shared memory.
// Replaces the reference's numpy.fft / mkl_fft calls (ipde/utilities.py:5-12).
#include "fft_generic.cuh"
#include "fft_short.cuh"
#include "fft_reg.cuh"
#include <algorithm>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------
// plans
static bool factorize(int n, std::vector<int>& fac) {
    fac.clear();
    if (n < 1) return false;
    // powers of two: radix 8 with the remainder as radix 4 (or 4*4 instead of 8*2) so that no pass is radix 2
    // unless n has a single factor of two -- every thread then owns 16 elements in every pass
    int a = 0;
    while (n % 2 == 0) { ++a; n /= 2; }
    while (a >= 3 && a != 4) { fac.push_back(8); a -= 3; }
    while (a >= 2) { fac.push_back(4); a -= 2; }
    if (a == 1) fac.push_back(2);
    const int odd[] = {3, 5, 7, 11, 13};
    for (int p : odd)
        while (n % p == 0) { fac.push_back(p); n /= p; }
    return n == 1 && (int)fac.size() <= IPDE_FFT_MAX_FACTORS;
}
// the same plus the remaining (large) prime factors: lengths the tile / generic kernels cannot do still get a plan;
// they run pass by pass through global memory with an O(R^2) butterfly (k_fft_pass_global_any) -- slow but exact,
// so a user grid of any size works as it does with numpy.fft
static bool factorize_any(int n, std::vector<int>& fac) {
    if (factorize(n, fac)) return true;
    int rest = n;
    for (int f : fac) rest /= f;
    for (int p = 17; (int64_t)p * p <= rest; p += 2)
        while (rest % p == 0) { fac.push_back(p); rest /= p; }
    if (rest > 1) fac.push_back(rest);
    return (int)fac.size() <= IPDE_FFT_MAX_FACTORS;
}

bool fft_size_supported(int n) {
    std::vector<int> f;
    return factorize(n, f);
}

int fft_next_size(int n) {
    if (n < 2) n = 2;
    if (n % 2) ++n;
    for (;; n += 2) {
        int m = n;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return n;
    }
}

static void fill_twiddles(std::vector<cplx>& t, int64_t n, int64_t count, int64_t denom) {
    t.resize(count);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (int64_t j = 0; j < count; ++j) {
        long double a = -tp * (long double)j / (long double)denom;
        t[j] = make_double2((double)cosl(a), (double)sinl(a));
    }
    (void)n;
}

int fft_plan_create(int n, FftPlan** out) {
    std::vector<int> fac;
    if (n < 1 || !factorize_any(n, fac)) {
        ipde_set_error("fft_plan_create: length %d is not supported", n);
        return IPDE_ERR_UNSUPPORTED;
    }
    FftPlan* p = new FftPlan();
    p->d.n = n;
    p->d.nfac = (int)fac.size();
    p->pow2 = p->d.nfac > 0;
    for (int i = 0; i < p->d.nfac; ++i) {
        p->d.fac[i] = fac[i];
        if (fac[i] > 5 && fac[i] != 8) p->fast = false;
        if (fac[i] != 4 && fac[i] != 8) p->pow2 = false;
        if (fac[i] > IPDE_FFT_MAX_RADIX) p->any = true;
    }
    std::vector<cplx> tw;
    fill_twiddles(tw, n, n, n);
    CUDA_TRY(cudaMalloc(&p->tw_dev, sizeof(cplx) * (size_t)n));
    CUDA_TRY(cudaMemcpy(p->tw_dev, tw.data(), sizeof(cplx) * (size_t)n, cudaMemcpyHostToDevice));
    fill_twiddles(tw, n, n + 1, 2 * (int64_t)n);
    CUDA_TRY(cudaMalloc(&p->half_tw_dev, sizeof(cplx) * (size_t)(n + 1)));
    CUDA_TRY(cudaMemcpy(p->half_tw_dev, tw.data(), sizeof(cplx) * (size_t)(n + 1), cudaMemcpyHostToDevice));
    p->d.tw = p->tw_dev;
    *out = p;
    return IPDE_OK;
}

void fft_plan_destroy(FftPlan* p) {
    if (!p) return;
    fft_plan_destroy(p->half);
    cudaFree(p->tw_dev);
    cudaFree(p->half_tw_dev);
    cudaFree(p->work);
    delete p;
}

static int ensure_work(FftPlan* p, size_t elems) {
    if (p->work_elems >= elems) return IPDE_OK;
    cudaFree(p->work);
    p->work = nullptr;
    p->work_elems = 0;
    CUDA_TRY(cudaMalloc(&p->work, sizeof(cplx) * elems));
    p->work_elems = elems;
    return IPDE_OK;
}

static inline int pow2_floor(int64_t x) {
    int p = 1;
    while ((int64_t)p * 2 <= x) p *= 2;
    return p;
}
static inline int ilog2(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}
// lines per tile for contiguous rows of length n: tiles of ~2048 elements (256 threads, 2-3 CTAs per SM)
static inline int rows_per_tile(int n, int64_t batch) {
    int c = (n >= 2048) ? 1 : pow2_floor(2048 / n);
    while (c > 1 && c > batch) c /= 2;
    return c;
}

// the cp.async double-buffered column kernel can be switched off (IPDE_FFT_PIPELINE=0) for A/B measurements
bool fft_pipeline_enabled();
static const bool g_fft_pipeline = []() { const char* e = getenv("IPDE_FFT_PIPELINE"); return !(e && e[0] == '0'); }();

bool fft_pipeline_enabled() { return g_fft_pipeline; }
static const bool g_fft_reg = []() { const char* e = getenv("IPDE_FFT_REG"); return !(e && e[0] == '0'); }();
bool fft_reg_enabled() { return g_fft_reg; }
template <int N>
static int launch_rows_reg(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
                           double scale, cudaStream_t st);


// elements per thread: 8 up to 4096-element tiles (<= 512 threads), 16 beyond
static inline int tile_ept(int64_t elems) { return elems > 4096 ? 16 : 8; }

// instantiate-and-launch over (elements per thread, radix class)
#define FFT_TILE_LAUNCH_ONE(K, E_, C_, g, t, smem, st, ...)                          \
    do {                                                                            \
        if ((t) <= 256) {                                                           \
            rc = set_smem(K<E_, C_, 256>, smem);                                    \
            if (!rc) K<E_, C_, 256><<<g, t, smem, st>>>(__VA_ARGS__);               \
        } else {                                                                    \
            rc = set_smem(K<E_, C_, 512>, smem);                                    \
            if (!rc) K<E_, C_, 512><<<g, t, smem, st>>>(__VA_ARGS__);               \
        }                                                                           \
    } while (0)
#define FFT_TILE_LAUNCH(K, ept, pow2, g, t, smem, st, ...)                                   \
    do {                                                                                    \
        if ((ept) == 8 && (pow2)) FFT_TILE_LAUNCH_ONE(K, 8, 0, g, t, smem, st, __VA_ARGS__); \
        else if ((ept) == 8) FFT_TILE_LAUNCH_ONE(K, 8, 1, g, t, smem, st, __VA_ARGS__);      \
        else if (pow2) FFT_TILE_LAUNCH_ONE(K, 16, 0, g, t, smem, st, __VA_ARGS__);           \
        else FFT_TILE_LAUNCH_ONE(K, 16, 1, g, t, smem, st, __VA_ARGS__);                     \
    } while (0)

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// contiguous rows, C lines per CTA
template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_fft_rows(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t batch, int64_t in_pitch,
           int64_t out_pitch, int C, int sign, double scale) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t l0 = (int64_t)blockIdx.x * C;
    const int cw = (int)min((int64_t)C, batch - l0);
    for (int c = 0; c < C; ++c) {
        if (c < cw) {
            const cplx* src = in + (l0 + c) * in_pitch;
            for (int e = tid; e < n; e += nt) smem[PADI(c * n + e)] = src[e];
        } else {
            for (int e = tid; e < n; e += nt) smem[PADI(c * n + e)] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    tile_fft<0, EPT, CLS>(smem, P, C, 0, sign, tid, nt);
    for (int c = 0; c < cw; ++c) {
        cplx* dst = out + (l0 + c) * out_pitch;
        for (int e = tid; e < n; e += nt) dst[e] = cscale(smem[PADI(c * n + e)], scale);
    }
}

// strided lines, C adjacent lines per CTA (C a power of two), optional four-step twiddle
template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_fft_cols(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, int64_t elem_stride, int64_t n_inner,
           int64_t outer_stride, int C, int logC, int sign, double scale, int tw_mode, const cplx* __restrict__ tw_big,
           RowWindow W) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t tiles = (n_inner + C - 1) >> logC;
    const int o = (int)(blockIdx.x / tiles);
    const int row_base = o * W.row_o;
    const int64_t c0 = (blockIdx.x - (int64_t)o * tiles) << logC;
    const int cw = (int)min((int64_t)C, n_inner - c0);
    const int64_t base = (int64_t)o * outer_stride + c0;
    const int total = n << logC;
    for (int i = tid; i < total; i += nt) {
        const int c = i & (C - 1), e = i >> logC;
        cplx v = make_double2(0.0, 0.0);
        const int row = row_base + e * W.row_e;
        if (c < cw && row >= W.in_lo && row < W.in_hi) {
            v = in[base + (int64_t)e * elem_stride + c];
            if (tw_mode == 2) {
                const cplx w = tw_big[o * e];
                v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
        }
        smem[PADI(i)] = v;
    }
    __syncthreads();
    tile_fft<1, EPT, CLS>(smem, P, C, logC, sign, tid, nt);
    for (int i = tid; i < total; i += nt) {
        const int c = i & (C - 1), e = i >> logC;
        const int row = row_base + e * W.row_e;
        if (c < cw && row >= W.out_lo && row < W.out_hi) {
            cplx v = smem[PADI(i)];
            if (tw_mode == 1) {
                const cplx w = tw_big[o * e];
                v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
            out[base + (int64_t)e * elem_stride + c] = cscale(v, scale);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Persistent, double-buffered version of k_fft_cols: a CTA walks over tiles with stride gridDim.x; while tile t is
// transformed in one shared-memory buffer, the 16-byte cp.async copies of tile t+1 are already in flight into the
// other one (zero-fill form for rows outside the input window), so the global-load latency is off the critical
// path regardless of how many CTAs fit an SM.
template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_fft_cols_pipe(FftPlanDev P, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t outer_stride, int C,
                int logC, int sign, double scale, int tw_mode, const cplx* __restrict__ tw_big, RowWindow W,
                int64_t ntiles_total, int buf_elems) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t tiles = (n_inner + C - 1) >> logC;
    const int total = n << logC;
    auto issue = [&](int64_t tile, cplx* buf) {
        const int o = (int)(tile / tiles);
        const int64_t c0 = (tile - (int64_t)o * tiles) << logC;
        const int cw = (int)min((int64_t)C, n_inner - c0);
        const int64_t base = (int64_t)o * outer_stride + c0;
        const int row_base = o * W.row_o;
        for (int i = tid; i < total; i += nt) {
            const int c = i & (C - 1), e = i >> logC;
            const int row = row_base + e * W.row_e;
            const bool live = (c < cw) && row >= W.in_lo && row < W.in_hi;
            cp_async16(buf + PADI(i), live ? in + base + (int64_t)e * elem_stride + c : in, live ? 16 : 0);
        }
        cp_async_commit();
    };
    int64_t tile = blockIdx.x;
    int cur = 0;
    if (tile < ntiles_total) issue(tile, smem);
    for (; tile < ntiles_total; tile += gridDim.x, cur ^= 1) {
        cplx* buf = smem + (size_t)cur * buf_elems;
        const int64_t next = tile + gridDim.x;
        if (next < ntiles_total) {
            issue(next, smem + (size_t)(cur ^ 1) * buf_elems);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int o = (int)(tile / tiles);
        const int64_t c0 = (tile - (int64_t)o * tiles) << logC;
        const int cw = (int)min((int64_t)C, n_inner - c0);
        const int64_t base = (int64_t)o * outer_stride + c0;
        const int row_base = o * W.row_o;
        if (tw_mode == 2) {
            for (int i = tid; i < total; i += nt) {
                const int e = i >> logC;
                const cplx w = tw_big[o * e];
                const cplx v = buf[PADI(i)];
                buf[PADI(i)] = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
            __syncthreads();
        }
        tile_fft<1, EPT, CLS>(buf, P, C, logC, sign, tid, nt);
        for (int i = tid; i < total; i += nt) {
            const int c = i & (C - 1), e = i >> logC;
            const int row = row_base + e * W.row_e;
            if (c < cw && row >= W.out_lo && row < W.out_hi) {
                cplx v = buf[PADI(i)];
                if (tw_mode == 1) {
                    const cplx w = tw_big[o * e];
                    v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
                }
                out[base + (int64_t)e * elem_stride + c] = cscale(v, scale);
            }
        }
        __syncthreads();     // the buffer is refilled two iterations from now by other threads' copies
    }
}

// ------------------------------------------------------------------------------------------
// Short power-of-two strided lines (N = RA*RB in {16, 32, 64}) held in registers, see fft_short.cuh.
// Same contract as k_fft_cols (four-step twiddles, row windows, in == out allowed: tiles are disjoint and a tile is
// read completely before it is written).
template <int RA, int RB, int SIGN, bool SCATTER>
__global__ void __launch_bounds__(128, 4)
k_fft_cols_short(const cplx* __restrict__ tw, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner,
                 int64_t outer_stride, double scale, int tw_mode, const cplx* __restrict__ tw_big, RowWindow W,
                 int tiles_per_outer, int ntiles_total, const SlabScatter sc) {
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
            const int e = short_elem_a<RA, RB>(jt, s);
            const int row = S.row_base + e * W.row_e;
            d[s] = (S.colok && row >= W.in_lo && row < W.in_hi) ? in[S.base + (int64_t)e * elem_stride]
                                                                : make_double2(0.0, 0.0);
        }
    };
    int tile = blockIdx.x;
    if (K::PREFETCH && tile < ntiles_total) load(tile, v);
    for (int buf = 0; tile < ntiles_total; tile += gridDim.x, buf ^= (K::NBUF - 1)) {
        if (K::PREFETCH) {
            const int next = tile + gridDim.x;
            if (next < ntiles_total) load(next, nx);
        } else {
            load(tile, v);
        }
        const ShortTile S = short_tile<K::LOGC>(tile, tiles_per_outer, c, n_inner, outer_stride, W.row_o);
        if (tw_mode == 2) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const cplx w = tw_big[S.o * short_elem_a<RA, RB>(jt, s)];
                v[s] = (SIGN < 0) ? cmul(v[s], w) : cmulc(v[s], w);
            }
        }
        short_fft<RA, RB, SIGN>(v, sm[buf], c, jt, tw);
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int e = short_elem_b<RA, RB>(jt, s);
            const int row = S.row_base + e * W.row_e;
            if (S.colok && row >= W.out_lo && row < W.out_hi) {
                cplx r = v[s];
                if (tw_mode == 1) {
                    const cplx w = tw_big[S.o * e];
                    r = (SIGN < 0) ? cmul(r, w) : cmulc(r, w);
                }
                if (SCATTER) {                            // backward slab transpose on the stores
                    if (sc.col0 + S.col < sc.P) *slab_bwd_addr(sc, row, S.col) = cscale(r, scale);
                } else {
                    out[S.base + (int64_t)e * elem_stride] = cscale(r, scale);
                }
            }
        }
        if (K::PREFETCH) {
#pragma unroll
            for (int s = 0; s < E; ++s) v[s] = nx[s];
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA variant of the plain short column pass (128- and 256-point lines, no row window): the tile of the NEXT
// iteration is fetched by one cp.async.bulk.tensor (3-D tensor map over {columns, line element, outer line}) into a
// second shared-memory buffer and signalled through an mbarrier, so no thread holds load addresses or waits on
// global loads; the registers only ever hold the tile being transformed.  ncu of the register-load version: 25 % warps
// active, long-scoreboard stalls 7-9 cycles per issued instruction (profiles/r02_ncu_k_fft_cols.txt).
#include <cuda.h>
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (TensorMapEncodeFn)f;
    }();
    return fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <int RA, int RB, int SIGN>
__global__ void __launch_bounds__(128, 3)
k_fft_cols_short_tma(const __grid_constant__ CUtensorMap tmap, const cplx* __restrict__ tw, cplx* out, int64_t elem_stride,
                     int64_t n_inner, int64_t outer_stride, double scale, int tw_mode, const cplx* __restrict__ tw_big,
                     int e_first, int tiles_per_outer, int ntiles_total) {
    using K = ShortCfg<RA, RB>;
    constexpr int E = K::EPT;
    extern __shared__ unsigned char smraw_[];
    // TMA destinations must be 128-byte aligned: align the dynamic segment by hand (128 spare bytes are requested)
    unsigned char* smraw = smraw_ + ((128u - (smem_u32(smraw_) & 127u)) & 127u);
    // two tile buffers; the buffer a tile arrived in doubles as the exchange buffer of its transform (the registers
    // hold the tile by then), which keeps a CTA at 64 KB and three CTAs on an SM
    cplx* tile_in[2] = {reinterpret_cast<cplx*>(smraw), reinterpret_cast<cplx*>(smraw) + K::TILE};
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile_in[1] + K::TILE);
    const int tid = threadIdx.x;
    const int c = tid & (K::C - 1), jt = tid >> K::LOGC;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tile, int b) {
        const int o = tile / tiles_per_outer;
        const int col0 = (tile - o * tiles_per_outer) << K::LOGC;
        mbar_expect_tx(&bar[b], (uint32_t)(K::TILE * sizeof(cplx)));
        // coordinates: (double index inside the row, second dimension, third dimension); e_first tells which of the
        // two strided dimensions is the line element (the map orders them by stride)
        if (e_first) tma_load_3d(tile_in[b], &tmap, 2 * col0, 0, o, &bar[b]);
        else tma_load_3d(tile_in[b], &tmap, 2 * col0, o, 0, &bar[b]);
    };
    int tile = blockIdx.x;
    uint32_t phase0 = 0, phase1 = 0;
    if (tid == 0 && tile < ntiles_total) issue(tile, 0);
    cplx v[E];
    for (int b = 0; tile < ntiles_total; tile += gridDim.x, b ^= 1) {
        const int next = tile + gridDim.x;
        // buffer b^1 served the previous iteration (tile and exchange); the barrier at the end of that iteration has
        // been passed by every thread
        if (tid == 0 && next < ntiles_total) issue(next, b ^ 1);
        if (b == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
        else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
        const ShortTile S = short_tile<K::LOGC>(tile, tiles_per_outer, c, n_inner, outer_stride, 0);
        const cplx* src = tile_in[b];
#pragma unroll
        for (int s = 0; s < E; ++s) v[s] = src[(short_elem_a<RA, RB>(jt, s) << K::LOGC) + c];
        if (tw_mode == 2) {
#pragma unroll
            for (int s = 0; s < E; ++s) {
                const cplx w = tw_big[S.o * short_elem_a<RA, RB>(jt, s)];
                v[s] = (SIGN < 0) ? cmul(v[s], w) : cmulc(v[s], w);
            }
        }
        short_fft<RA, RB, SIGN>(v, tile_in[b], c, jt, tw);     // NBUF == 1: starts with a barrier (tile reads done)
        __syncthreads();                                       // exchange reads done: the buffer may be refilled
#pragma unroll
        for (int s = 0; s < E; ++s) {
            const int e = short_elem_b<RA, RB>(jt, s);
            if (S.colok) {
                cplx r = v[s];
                if (tw_mode == 1) {
                    const cplx w = tw_big[S.o * e];
                    r = (SIGN < 0) ? cmul(r, w) : cmulc(r, w);
                }
                out[S.base + (int64_t)e * elem_stride] = cscale(r, scale);
            }
        }
    }
}

// Measured on B200 (16384^2 4-boundary solve, 8 launches of the 128/256-point passes, CUDA events): register loads
// 19.7 ms (3.51 TB/s); TMA, 2 CTAs/SM with a separate exchange buffer 21.1 ms; TMA, 3 CTAs/SM with the tile buffer
// reused for the exchange 21.8 ms.  The pass is limited by its strided 128-256-byte stores and the FP64 butterflies,
// not by load latency, so the register-load kernel stays the default; IPDE_FFT_TMA=1 selects this one.
static const bool g_fft_tma = []() { const char* e = getenv("IPDE_FFT_TMA"); return e && e[0] == '1'; }();

template <int RA, int RB>
static int launch_cols_short_tma(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                                 int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big,
                                 cudaStream_t st, bool* done) {
    using K = ShortCfg<RA, RB>;
    *done = false;
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return IPDE_OK;
    // tensor map over doubles: {2*n_inner, then the two strided dimensions by increasing stride}
    const bool e_first = elem_stride <= outer_stride || n_outer == 1;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    dims[0] = (cuuint64_t)(2 * n_inner);
    box[0] = (cuuint32_t)(2 * K::C);
    if (e_first) {
        dims[1] = K::N; dims[2] = (cuuint64_t)n_outer;
        strides[0] = (cuuint64_t)elem_stride * 16; strides[1] = (cuuint64_t)(n_outer == 1 ? elem_stride * K::N : outer_stride) * 16;
        box[1] = K::N; box[2] = 1;
    } else {
        dims[1] = (cuuint64_t)n_outer; dims[2] = K::N;
        strides[0] = (cuuint64_t)outer_stride * 16; strides[1] = (cuuint64_t)elem_stride * 16;
        box[1] = 1; box[2] = K::N;
    }
    CUtensorMap map;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<cplx*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return IPDE_OK;          // shapes the descriptor cannot express: register-load kernel
    const int64_t tiles = cdiv(n_inner, (int64_t)K::C);
    const int64_t total = tiles * n_outer;
    ARG_CHECK(total < (int64_t)1 << 30, "short column transform: too many tiles");
    const size_t smem = 2 * (size_t)K::TILE * sizeof(cplx) + 128 + 64;
    const unsigned g = (unsigned)std::min<int64_t>(total, (int64_t)ipde_num_sms() * 3);
    int rc;
    if (sign < 0) {
        rc = set_smem(k_fft_cols_short_tma<RA, RB, -1>, smem);
        if (rc) return rc;
        k_fft_cols_short_tma<RA, RB, -1><<<g, 128, smem, st>>>(map, p->d.tw, out, elem_stride, n_inner, outer_stride, scale,
                                                               tw_mode, tw_big, e_first ? 1 : 0, (int)tiles, (int)total);
    } else {
        rc = set_smem(k_fft_cols_short_tma<RA, RB, +1>, smem);
        if (rc) return rc;
        k_fft_cols_short_tma<RA, RB, +1><<<g, 128, smem, st>>>(map, p->d.tw, out, elem_stride, n_inner, outer_stride, scale,
                                                               tw_mode, tw_big, e_first ? 1 : 0, (int)tiles, (int)total);
    }
    KERNEL_CHECK();
    *done = true;
    return IPDE_OK;
}

static const bool g_fft_short = []() { const char* e = getenv("IPDE_FFT_SHORT"); return !(e && e[0] == '0'); }();
bool fft_short_enabled() { return g_fft_short; }

template <int RA, int RB>
static int launch_cols_short(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                             int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, RowWindow W,
                             cudaStream_t st, const SlabScatter& sc) {
    using K = ShortCfg<RA, RB>;
    const int64_t tiles = cdiv(n_inner, (int64_t)K::C);
    const int64_t total = tiles * n_outer;
    ARG_CHECK(total < (int64_t)1 << 30, "short column transform: too many tiles");
    const int64_t per_sm = 4;
    const unsigned g = (unsigned)std::min<int64_t>(total, (int64_t)ipde_num_sms() * per_sm);
    if (sc.R) {
        ARG_CHECK(sign > 0, "short column transform: scattered stores are built for the inverse pass");
        k_fft_cols_short<RA, RB, +1, true><<<g, 128, 0, st>>>(p->d.tw, in, out, elem_stride, n_inner, outer_stride, scale,
                                                              tw_mode, tw_big, W, (int)tiles, (int)total, sc);
    } else if (sign < 0)
        k_fft_cols_short<RA, RB, -1, false><<<g, 128, 0, st>>>(p->d.tw, in, out, elem_stride, n_inner, outer_stride, scale,
                                                               tw_mode, tw_big, W, (int)tiles, (int)total, sc);
    else
        k_fft_cols_short<RA, RB, +1, false><<<g, 128, 0, st>>>(p->d.tw, in, out, elem_stride, n_inner, outer_stride, scale,
                                                               tw_mode, tw_big, W, (int)tiles, (int)total, sc);
    KERNEL_CHECK();
    return IPDE_OK;
}

// global-memory Stockham pass (any line length): one work item per butterfly
template <int R>
__global__ void k_fft_pass_global(FftPlanDev P, const cplx* __restrict__ src, cplx* __restrict__ dst, int64_t batch,
                                  int64_t src_line, int64_t src_elem, int64_t dst_line, int64_t dst_elem, int Ns,
                                  int sign, double scale) {
    const int n = P.n;
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const int64_t total = batch * m;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t line = it / m;
        const int j = (int)(it % m);
        const int k = j % Ns;
        const cplx* s = src + line * src_line;
        cplx* d = dst + line * dst_line;
        cplx v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = s[(int64_t)(j + t * m) * src_elem];
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
        for (int t = 0; t < R; ++t) d[(int64_t)(j0 + t * Ns) * dst_elem] = cscale(v[t], scale);
    }
}

__global__ void k_gather_lines(const cplx* __restrict__ src, cplx* __restrict__ dst, int64_t batch, int n, int64_t src_line,
                               int64_t src_elem) {
    const int64_t total = batch * n;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x)
        dst[it] = src[(it / n) * src_line + (it % n) * src_elem];
}

// the same pass for an arbitrary (large prime) radix: one work item per OUTPUT, an O(R) sum each
__global__ void k_fft_pass_global_any(FftPlanDev P, const cplx* __restrict__ src, cplx* __restrict__ dst, int64_t batch,
                                      int64_t src_line, int64_t src_elem, int64_t dst_line, int64_t dst_elem, int Ns, int R,
                                      int sign, double scale) {
    const int n = P.n;
    const int m = n / R;
    const int tstep = n / (Ns * R);
    const int64_t total = batch * n;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t line = it / n;
        const int rem = (int)(it % n);
        const int j = rem % m, t = rem / m;
        const int k = j % Ns;
        const cplx* s = src + line * src_line;
        cplx acc = make_double2(0.0, 0.0);
        for (int q = 0; q < R; ++q) {
            // twiddle w_{Ns R}^{k q} times the butterfly root w_R^{q t}, both as powers of w_n
            const int64_t e = ((int64_t)q * k * tstep + (int64_t)(((int64_t)q * t) % R) * m) % n;
            const cplx w = P.tw[e];
            const cplx x = s[(int64_t)(j + q * m) * src_elem];
            acc = cadd(acc, (sign < 0) ? cmul(x, w) : cmulc(x, w));
        }
        dst[line * dst_line + (int64_t)((j - k) * R + k + t * Ns) * dst_elem] = cscale(acc, scale);
    }
}

// Lines too long for shared memory: pass-per-launch through `work` (contiguous lines of n).
static int fft_lines_global(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_line, int64_t in_elem,
                            int64_t out_line, int64_t out_elem, int sign, double scale, cudaStream_t st) {
    const int n = p->d.n;
    int rc = ensure_work(p, (size_t)2 * n * batch);
    if (rc) return rc;
    cplx* w0 = p->work;
    cplx* w1 = p->work + (size_t)n * batch;
    const cplx* src = in;
    int64_t sl = in_line, se = in_elem;
    int Ns = 1;
    for (int f = 0; f < p->d.nfac; ++f) {
        const int R = p->d.fac[f];
        const bool last = (f == p->d.nfac - 1);
        cplx* dst = last ? out : ((f % 2 == 0) ? w0 : w1);
        int64_t dl = last ? out_line : n, de = last ? out_elem : 1;
        const double sc = last ? scale : 1.0;
        const int64_t total = batch * (n / R);
        const int threads = 256;
        const int blocks = (int)std::min<int64_t>(cdiv(total, threads), (int64_t)ipde_num_sms() * 16);
#define LAUNCH(RR) k_fft_pass_global<RR><<<blocks, threads, 0, st>>>(p->d, src, dst, batch, sl, se, dl, de, Ns, sign, sc)
        switch (R) {
            case 2: LAUNCH(2); break;
            case 3: LAUNCH(3); break;
            case 4: LAUNCH(4); break;
            case 5: LAUNCH(5); break;
            case 7: LAUNCH(7); break;
            case 8: LAUNCH(8); break;
            case 11: LAUNCH(11); break;
            case 13: LAUNCH(13); break;
            default: {
                // every output reads the whole butterfly: an in-place single pass (prime n) goes through a copy
                if (src == dst) {
                    const int64_t tot = batch * n;
                    const int b3 = (int)std::min<int64_t>(cdiv(tot, threads), (int64_t)ipde_num_sms() * 16);
                    k_gather_lines<<<b3, threads, 0, st>>>(src, w1, batch, n, sl, se);
                    KERNEL_CHECK();
                    src = w1; sl = n; se = 1;
                }
                const int64_t outs = batch * n;
                const int b2 = (int)std::min<int64_t>(cdiv(outs, threads), (int64_t)ipde_num_sms() * 16);
                k_fft_pass_global_any<<<b2, threads, 0, st>>>(p->d, src, dst, batch, sl, se, dl, de, Ns, R, sign, sc);
            }
        }
#undef LAUNCH
        KERNEL_CHECK();
        src = dst; sl = dl; se = de;
        Ns *= R;
    }
    if (p->d.nfac == 0 && out != in) {   // n == 1
        CUDA_TRY(cudaMemcpyAsync(out, in, sizeof(cplx) * batch, cudaMemcpyDeviceToDevice, st));
    }
    return IPDE_OK;
}

static int fft_split_radix(int n);
static int fft_rows_split(FftPlan* p, int R, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                          int sign, double scale, cudaStream_t st);

int fft_rows(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
             double scale, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (g_fft_reg && p->fast) {
        const int R = fft_split_radix(n);
        if (R) return fft_rows_split(p, R, in, out, batch, in_pitch, out_pitch, sign, scale, st);
    }
    if (!p->fast && !p->any && n <= IPDE_FFT_GENERIC_MAX) return fft_rows_generic(p, in, out, batch, in_pitch, out_pitch, sign, scale, st);
    if (n > IPDE_FFT_SMEM_MAX || !p->fast) return fft_lines_global(p, in, out, batch, in_pitch, 1, out_pitch, 1, sign, scale, st);
    const int C = rows_per_tile(n, batch);
    const size_t smem = fft_tile_smem((int64_t)n * C);
    const int ept = tile_ept((int64_t)n * C);
    const unsigned g = (unsigned)cdiv(batch, C);
    const int t = fft_tile_threads(n, C, ept);
    int rc = IPDE_OK;
    KERNEL_TAG("k_fft_rows", 32.0 * (double)n * (double)batch);
    if (g_fft_reg && fft_reg_size(n)) {
#define CALL_REG(N_) return launch_rows_reg<N_>(p, in, out, batch, in_pitch, out_pitch, sign, scale, st)
        REG_DISPATCH(n, CALL_REG);
#undef CALL_REG
    }
    FFT_TILE_LAUNCH(k_fft_rows, ept, p->pow2, g, t, smem, st, p->d, in, out, batch, in_pitch, out_pitch, C, sign, scale);
    if (rc) return rc;
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_cols(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
             int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, int64_t n_big,
             cudaStream_t st) {
    if (n_inner <= 0 || n_outer <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (p->fast && n <= IPDE_FFT_SMEM_MAX)
        return fft_cols_windowed(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big,
                                 row_window_all(), st);
    if (!p->fast && !p->any && n <= IPDE_FFT_GENERIC_MAX)
        return fft_cols_generic(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big, n_big, st);
    if (n > IPDE_FFT_SMEM_MAX || !p->fast) {
        if (tw_mode != 0) {
            ipde_set_error("fft_cols: twiddled transform of length %d does not fit shared memory", n);
            return IPDE_ERR_UNSUPPORTED;
        }
        // treat every (outer, inner) pair as its own strided line
        for (int64_t o = 0; o < n_outer; ++o) {
            int rc = fft_lines_global(p, in + o * outer_stride, out + o * outer_stride, n_inner, 1, elem_stride, 1,
                                      elem_stride, sign, scale, st);
            if (rc) return rc;
        }
        return IPDE_OK;
    }
    ipde_set_error("fft_cols: unreachable");
    return IPDE_ERR_UNSUPPORTED;
}

int fft_cols_windowed(FftPlan* p, const cplx* in, cplx* out, int64_t elem_stride, int64_t n_inner, int64_t n_outer,
                      int64_t outer_stride, int sign, double scale, int tw_mode, const cplx* tw_big, RowWindow W,
                      cudaStream_t st, const SlabScatter* sc) {
    if (n_inner <= 0 || n_outer <= 0) return IPDE_OK;
    const int n = p->d.n;
    ARG_CHECK(p->fast && n <= IPDE_FFT_SMEM_MAX, "windowed column transform needs a tile-engine length");
    if (sc && sc->R) {
        // slab transpose on the stores: only the register-resident short pass implements it
        if (!(g_fft_short && p->pow2 && fft_short_size_plain(n) && W.row_e != 0)) {
            ipde_set_error("fft_cols: scattered stores need a short power-of-two four-step pass (n = %d)", n);
            return IPDE_ERR_UNSUPPORTED;
        }
        KERNEL_TAG("k_fft_cols", 32.0 * (double)n * (double)n_inner * (double)n_outer);
#define CALL_SHORT(RA_, RB_) return launch_cols_short<RA_, RB_>(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big, W, st, *sc)
        SHORT_DISPATCH_PLAIN(n, CALL_SHORT);
#undef CALL_SHORT
    }
    int C = pow2_floor(IPDE_FFT_SMEM_MAX / n);
    if (C > 16) C = 16;
    while (C > 1 && C / 2 >= n_inner) C /= 2;
    const int logC = ilog2(C);
    const size_t smem = fft_tile_smem((int64_t)n * C);
    const int ept = tile_ept((int64_t)n * C);
    const int64_t tiles = cdiv(n_inner, C);
    const unsigned g = (unsigned)(tiles * n_outer);
    const int t = fft_tile_threads(n, C, ept);
    int rc = IPDE_OK;
    // algorithmic bytes: rows actually read + rows actually written
    double fin = 1.0, fout = 1.0;
    if (W.row_e != 0 || W.row_o != 0) {
        const double rows = (double)n * (double)n_outer;
        fin = std::min(1.0, std::max(0.0, (double)(W.in_hi - W.in_lo)) / rows);
        fout = std::min(1.0, std::max(0.0, (double)(W.out_hi - W.out_lo)) / rows);
    }
    KERNEL_TAG("k_fft_cols", 16.0 * (fin + fout) * (double)n * (double)n_inner * (double)n_outer);
    if (g_fft_short && g_fft_tma && p->pow2 && (n == 128 || n == 256) && W.row_e == 0 && W.row_o == 0 &&
        (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (elem_stride * 16) % 16 == 0) {
        bool done = false;
        int rc_t;
        if (n == 128) rc_t = launch_cols_short_tma<16, 8>(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big, st, &done);
        else rc_t = launch_cols_short_tma<16, 16>(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big, st, &done);
        if (rc_t) return rc_t;
        if (done) return IPDE_OK;
    }
    if (g_fft_short && p->pow2 && fft_short_size_plain(n)) {
#define CALL_SHORT(RA_, RB_) return launch_cols_short<RA_, RB_>(p, in, out, elem_stride, n_inner, n_outer, outer_stride, sign, scale, tw_mode, tw_big, W, st, slab_scatter_none())
        SHORT_DISPATCH_PLAIN(n, CALL_SHORT);
#undef CALL_SHORT
    }
    const int64_t ntiles_total = (int64_t)tiles * n_outer;
    const int buf_elems = (int)(PADI((int64_t)n * C) + 1);
    const size_t smem2 = 2 * sizeof(cplx) * (size_t)buf_elems;
    const int per_sm = std::max<int>(1, std::min<int>(4, (int)((200 * 1024) / smem2)));
    const int64_t pgrid = (int64_t)ipde_num_sms() * per_sm;
    if (g_fft_pipeline && smem2 <= 200 * 1024 && ntiles_total >= 2 * pgrid) {
        const unsigned gp = (unsigned)pgrid;
        FFT_TILE_LAUNCH(k_fft_cols_pipe, ept, p->pow2, gp, t, smem2, st, p->d, in, out, elem_stride, n_inner, outer_stride, C, logC,
                        sign, scale, tw_mode, tw_big, W, ntiles_total, buf_elems);
    } else {
        FFT_TILE_LAUNCH(k_fft_cols, ept, p->pow2, g, t, smem, st, p->d, in, out, elem_stride, n_inner, outer_stride, C, logC, sign, scale, tw_mode, tw_big, W);
    }
    if (rc) return rc;
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// real rows (C lines per CTA): packed complex FFT of length n + untangling, fused
template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_r2c_rows(FftPlanDev P, const cplx* __restrict__ half_tw, const double* __restrict__ in, cplx* __restrict__ out,
           int64_t batch, int64_t in_pitch, int64_t out_pitch, int C) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;   // n = complex length = real length / 2
    const int64_t l0 = (int64_t)blockIdx.x * C;
    const int cw = (int)min((int64_t)C, batch - l0);
    for (int c = 0; c < C; ++c) {
        if (c < cw) {
            const cplx* src = reinterpret_cast<const cplx*>(in + (l0 + c) * in_pitch);
            for (int e = tid; e < n; e += nt) smem[PADI(c * n + e)] = src[e];
        } else {
            for (int e = tid; e < n; e += nt) smem[PADI(c * n + e)] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    tile_fft<0, EPT, CLS>(smem, P, C, 0, -1, tid, nt);
    for (int c = 0; c < cw; ++c) {
        cplx* dst = out + (l0 + c) * out_pitch;
        const int b = c * n;
        for (int k = tid; k <= n; k += nt) {
            const cplx zk = smem[PADI(b + (k == n ? 0 : k))];
            const cplx zc = cconj(smem[PADI(b + (k == 0 ? 0 : n - k))]);
            const cplx E = cscale(cadd(zk, zc), 0.5);
            const cplx O = cmulni(cscale(csub(zk, zc), 0.5));
            dst[k] = cadd(E, cmul(half_tw[k], O));
        }
    }
}

template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_c2r_rows(FftPlanDev P, const cplx* __restrict__ half_tw, const cplx* __restrict__ in, double* __restrict__ out,
           int64_t batch, int64_t in_pitch, int64_t out_pitch, int C, double scale) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t l0 = (int64_t)blockIdx.x * C;
    const int cw = (int)min((int64_t)C, batch - l0);
    for (int c = 0; c < C; ++c) {
        const int b = c * n;
        if (c < cw) {
            const cplx* src = in + (l0 + c) * in_pitch;
            for (int k = tid; k < n; k += nt) {
                const cplx xk = src[k];
                const cplx xc = cconj(src[n - k]);
                const cplx E = cscale(cadd(xk, xc), 0.5);
                const cplx O = cmulc(cscale(csub(xk, xc), 0.5), half_tw[k]);
                smem[PADI(b + k)] = cadd(E, cmuli(O));
            }
        } else {
            for (int k = tid; k < n; k += nt) smem[PADI(b + k)] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    tile_fft<0, EPT, CLS>(smem, P, C, 0, +1, tid, nt);
    for (int c = 0; c < cw; ++c) {
        cplx* dst = reinterpret_cast<cplx*>(out + (l0 + c) * out_pitch);
        for (int e = tid; e < n; e += nt) dst[e] = cscale(smem[PADI(c * n + e)], scale);
    }
}

// ------------------------------------------------------------------------------------------
// Real rows through the register-resident line transform (fft_reg.cuh).  One tile of L lines per CTA.
// r2c: packed complex transform, then the untangling step needs Z[k] and Z[n-k]: one more trip through the line
// buffer.  c2r: the tangling step reads X[k] and X[n-k] straight from global memory (the mirrored read is
// coalesced too and hits L1/L2), so the transform starts in registers.

template <int N>
__global__ void __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_r2c_rows_reg(const cplx* __restrict__ tw, const cplx* __restrict__ half_tw, const double* __restrict__ in,
               cplx* __restrict__ out, int64_t batch, int64_t in_pitch, int64_t out_pitch, const SlabScatter sc) {
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    extern __shared__ cplx smem[];
    const int tid = threadIdx.x;
    const int l = tid >> G::LOGT, jt = tid & (G::T - 1);
    cplx* line = smem + l * G::LINE;
    const int64_t row = (int64_t)blockIdx.x * G::L + l;
    const bool live = row < batch;
    cplx v[G::EPT];
    const cplx* src = reinterpret_cast<const cplx*>(in + (live ? row : 0) * in_pitch);
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) v[s] = live ? src[reg_elem<N, K::R0>(jt, s)] : make_double2(0.0, 0.0);
    reg_fft<N, -1, false>(v, line, jt, tw);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) line[reg_pad<G::SH>(reg_elem<N, G::RLAST>(jt, s))] = v[s];
    __syncthreads();
    if (live) {
        cplx* dst = out + row * out_pitch;
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) {
            const int k = reg_elem<N, G::RLAST>(jt, s);
            const cplx zk = v[s];
            const cplx zc = cconj(line[reg_pad<G::SH>(k == 0 ? 0 : N - k)]);
            const cplx E = cscale(cadd(zk, zc), 0.5);
            const cplx O = cmulni(cscale(csub(zk, zc), 0.5));
            if (sc.R) {                                   // forward slab transpose on the stores
                *slab_fwd_addr(sc, row, k) = cadd(E, cmul(half_tw[k], O));
                if (k == 0) *slab_fwd_addr(sc, row, N) = cadd(E, cmul(half_tw[N], O));
            } else {
                dst[k] = cadd(E, cmul(half_tw[k], O));
                if (k == 0) dst[N] = cadd(E, cmul(half_tw[N], O));
            }
        }
    }
}

template <int N>
__global__ void __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_c2r_rows_reg(const cplx* __restrict__ tw, const cplx* __restrict__ half_tw, const cplx* __restrict__ in,
               double* __restrict__ out, int64_t batch, int64_t in_pitch, int64_t out_pitch, double scale,
               const uint32_t* __restrict__ mask, int64_t mask_pitch) {
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    extern __shared__ cplx smem[];
    const int tid = threadIdx.x;
    const int l = tid >> G::LOGT, jt = tid & (G::T - 1);
    cplx* line = smem + l * G::LINE;
    const int64_t row = (int64_t)blockIdx.x * G::L + l;
    const bool live = row < batch;
    cplx v[G::EPT];
    const cplx* src = in + (live ? row : 0) * in_pitch;
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) {
        const int k = reg_elem<N, K::R0>(jt, s);
        if (live) {
            const cplx xk = src[k];
            const cplx xc = cconj(src[N - k]);
            const cplx E = cscale(cadd(xk, xc), 0.5);
            const cplx O = cmulc(cscale(csub(xk, xc), 0.5), half_tw[k]);
            v[s] = cadd(E, cmuli(O));
        } else {
            v[s] = make_double2(0.0, 0.0);
        }
    }
    reg_fft<N, +1, false>(v, line, jt, tw);
    if (live) {
        cplx* dst = reinterpret_cast<cplx*>(out + row * out_pitch);
        if (mask) {                                    // bit per block of 64 reals = 32 packed samples
            const uint32_t* mrow = mask + row * mask_pitch;
#pragma unroll
            for (int s = 0; s < G::EPT; ++s) {
                const int e = reg_elem<N, G::RLAST>(jt, s);
                if ((mrow[e >> 10] >> ((e >> 5) & 31)) & 1u) dst[e] = cscale(v[s], scale);
            }
        } else {
#pragma unroll
            for (int s = 0; s < G::EPT; ++s) dst[reg_elem<N, G::RLAST>(jt, s)] = cscale(v[s], scale);
        }
    }
}

template <int N>
static int launch_r2c_reg(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                          cudaStream_t st, const SlabScatter& sc = slab_scatter_none()) {
    using G = RegGeom<N>;
    int rc = set_smem(k_r2c_rows_reg<N>, G::SMEM);
    if (rc) return rc;
    k_r2c_rows_reg<N><<<(unsigned)cdiv(batch, (int64_t)G::L), G::NT, G::SMEM, st>>>(p->d.tw, p->half_tw_dev, in, out, batch,
                                                                                   in_pitch, out_pitch, sc);
    KERNEL_CHECK();
    return IPDE_OK;
}
template <int N>
static int launch_c2r_reg(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                          double scale, const uint32_t* mask, int64_t mask_pitch, cudaStream_t st) {
    using G = RegGeom<N>;
    int rc = set_smem(k_c2r_rows_reg<N>, G::SMEM);
    if (rc) return rc;
    k_c2r_rows_reg<N><<<(unsigned)cdiv(batch, (int64_t)G::L), G::NT, G::SMEM, st>>>(p->d.tw, p->half_tw_dev, in, out, batch,
                                                                                   in_pitch, out_pitch, scale, mask, mask_pitch);
    KERNEL_CHECK();
    return IPDE_OK;
}

// Complex rows and the annular variable-coefficient multiply through the register-resident line transform.
template <int N, int SIGN>
__global__ void __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_fft_rows_reg(const cplx* __restrict__ tw, const cplx* in, cplx* out, int64_t batch,
               int64_t in_pitch, int64_t out_pitch, double scale) {
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    extern __shared__ cplx smem[];
    const int tid = threadIdx.x;
    const int l = tid >> G::LOGT, jt = tid & (G::T - 1);
    cplx* line = smem + l * G::LINE;
    const int64_t row = (int64_t)blockIdx.x * G::L + l;
    const bool live = row < batch;
    cplx v[G::EPT];
    const cplx* src = in + (live ? row : 0) * in_pitch;
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) v[s] = live ? src[reg_elem<N, K::R0>(jt, s)] : make_double2(0.0, 0.0);
    reg_fft<N, SIGN, false>(v, line, jt, tw);
    if (live) {
        cplx* dst = out + row * out_pitch;
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) dst[reg_elem<N, G::RLAST>(jt, s)] = cscale(v[s], scale);
    }
}

// (launched with a handful of lines from the GMRES operator: latency matters, occupancy does not -- no register cap)
template <int N>
__global__ void __launch_bounds__(RegGeom<N>::NT, 1)
k_fft_mul_fft_rows_reg(const cplx* __restrict__ tw, const cplx* in, cplx* out,
                       const double* __restrict__ m, int64_t batch) {
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    extern __shared__ cplx smem[];
    const int tid = threadIdx.x;
    const int l = tid >> G::LOGT, jt = tid & (G::T - 1);
    cplx* line = smem + l * G::LINE;
    const int64_t row = (int64_t)blockIdx.x * G::L + l;
    const bool live = row < batch;
    const int64_t off = (live ? row : 0) * N;
    cplx v[G::EPT];
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) v[s] = live ? in[off + reg_elem<N, K::R0>(jt, s)] : make_double2(0.0, 0.0);
    reg_fft<N, +1, false>(v, line, jt, tw);
    const double inv = 1.0 / N;
    if constexpr (G::RLAST == K::R0) {
        // the outputs of the last stage sit exactly where the first stage of the next transform wants its inputs
#pragma unroll
        for (int s = 0; s < G::EPT; ++s)
            v[s] = cscale(v[s], live ? m[off + reg_elem<N, K::R0>(jt, s)] * inv : 0.0);
    } else {
        __syncthreads();
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) {
            const int e = reg_elem<N, G::RLAST>(jt, s);
            line[reg_pad<G::SH>(e)] = cscale(v[s], live ? m[off + e] * inv : 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) v[s] = line[reg_pad<G::SH>(reg_elem<N, K::R0>(jt, s))];
    }
    reg_fft<N, -1, true>(v, line, jt, tw);
    if (live) {
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) out[off + reg_elem<N, G::RLAST>(jt, s)] = v[s];
    }
}

template <int N>
static int launch_rows_reg(FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign,
                           double scale, cudaStream_t st) {
    using G = RegGeom<N>;
    const unsigned g = (unsigned)cdiv(batch, (int64_t)G::L);
    int rc;
    if (sign < 0) {
        rc = set_smem(k_fft_rows_reg<N, -1>, G::SMEM);
        if (rc) return rc;
        k_fft_rows_reg<N, -1><<<g, G::NT, G::SMEM, st>>>(p->d.tw, in, out, batch, in_pitch, out_pitch, scale);
    } else {
        rc = set_smem(k_fft_rows_reg<N, +1>, G::SMEM);
        if (rc) return rc;
        k_fft_rows_reg<N, +1><<<g, G::NT, G::SMEM, st>>>(p->d.tw, in, out, batch, in_pitch, out_pitch, scale);
    }
    KERNEL_CHECK();
    return IPDE_OK;
}
template <int N>
static int launch_mul_fft_reg(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st) {
    using G = RegGeom<N>;
    int rc = set_smem(k_fft_mul_fft_rows_reg<N>, G::SMEM);
    if (rc) return rc;
    k_fft_mul_fft_rows_reg<N><<<(unsigned)cdiv(batch, (int64_t)G::L), G::NT, G::SMEM, st>>>(p->d.tw, in, out, m, batch);
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// Long contiguous lines, n = 2m with m a register-kernel length (n = 16384: the rows of the 32768-point boxes of the
// free-space evaluator and of the NUFFT fine grid at 16384^2).  One radix-2 step around the m-point register
// transform instead of five global-memory passes:
//   forward  (decimation in time):  E_r = FFT_m(z[2j + r]), r = 0, 1  -- the m-point kernel reads its line with an
//            element stride of 2 (both sub-lines of a row are read by neighbouring CTAs, the sectors are shared
//            through L2) --, then Z[k] = E_0[k] + w^k E_1[k], Z[k + m] = E_0[k] - w^k E_1[k] in the consumer;
//   inverse  (decimation in frequency):  A[k] = Z[k] + Z[k+m], B[k] = (Z[k] - Z[k+m]) conj(w^k) in the producer,
//            then z[2j] = IFFT_m(A)[j], z[2j+1] = IFFT_m(B)[j]: the m-point kernel stores with an element stride 2.
template <int N, int SIGN>
__global__ void __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_fft_rows_reg_sub(const cplx* __restrict__ tw, const cplx* in, cplx* out, int64_t nsub, int R, int64_t in_pitch, int64_t in_sub,
                   int in_es, int64_t out_pitch, int64_t out_sub, int out_es, double scale, const double* __restrict__ mul,
                   int64_t mul_pitch, int64_t in_nz) {
    // in_nz: input elements with (strided) index >= in_nz are zero and are not read
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    extern __shared__ cplx smem[];
    const int tid = threadIdx.x;
    const int l = tid >> G::LOGT, jt = tid & (G::T - 1);
    cplx* line = smem + l * G::LINE;
    const int64_t sub = (int64_t)blockIdx.x * G::L + l;          // sub-line index: R*row + r
    const bool live = sub < nsub;
    const int64_t row = (live ? sub : 0) / R;
    const int r = (int)((live ? sub : 0) - row * R);
    cplx v[G::EPT];
    const cplx* src = in + row * in_pitch + r * in_sub;
    if (mul) {
        const double* mm = mul + row * mul_pitch + r * in_sub;
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) {
            const int64_t a = (int64_t)reg_elem<N, K::R0>(jt, s) * in_es;
            v[s] = live ? cscale(src[a], mm[a]) : make_double2(0.0, 0.0);
        }
    } else {
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) {
            const int64_t a = (int64_t)reg_elem<N, K::R0>(jt, s) * in_es;
            v[s] = (live && a + r * in_sub < in_nz) ? src[a] : make_double2(0.0, 0.0);
        }
    }
    reg_fft<N, SIGN, false>(v, line, jt, tw);
    if (live) {
        cplx* dst = out + row * out_pitch + r * out_sub;
#pragma unroll
        for (int s = 0; s < G::EPT; ++s) dst[(int64_t)reg_elem<N, G::RLAST>(jt, s) * out_es] = cscale(v[s], scale);
    }
}

template <int N>
static int launch_rows_reg_sub(FftPlan* ph, const cplx* in, cplx* out, int64_t nsub, int R, int64_t in_pitch, int64_t in_sub, int in_es,
                               int64_t out_pitch, int64_t out_sub, int out_es, int sign, double scale, const double* mul,
                               int64_t mul_pitch, int64_t in_nz, cudaStream_t st) {
    using G = RegGeom<N>;
    const unsigned g = (unsigned)cdiv(nsub, (int64_t)G::L);
    int rc;
    if (sign < 0) {
        rc = set_smem(k_fft_rows_reg_sub<N, -1>, G::SMEM);
        if (rc) return rc;
        k_fft_rows_reg_sub<N, -1><<<g, G::NT, G::SMEM, st>>>(ph->d.tw, in, out, nsub, R, in_pitch, in_sub, in_es, out_pitch,
                                                             out_sub, out_es, scale, mul, mul_pitch, in_nz);
    } else {
        rc = set_smem(k_fft_rows_reg_sub<N, +1>, G::SMEM);
        if (rc) return rc;
        k_fft_rows_reg_sub<N, +1><<<g, G::NT, G::SMEM, st>>>(ph->d.tw, in, out, nsub, R, in_pitch, in_sub, in_es, out_pitch,
                                                             out_sub, out_es, scale, mul, mul_pitch, in_nz);
    }
    KERNEL_CHECK();
    return IPDE_OK;
}

// c2r of rows of 2n reals, n = 2m: the tangle step (half spectrum X[0..n] -> packed spectrum Z) and the radix-2
// split (A = Z[k] + Z[k+m], B = (Z[k] - Z[k+m]) conj(w^k)) are applied on the loads of the m-point inverse transform.
// A row is owned by a CLUSTER OF TWO CTAs: rank 0 transforms A (even packed samples), rank 1 transforms B (odd ones).
//   loads    rank c tangles and splits the half k in [c m/2, (c+1) m/2) of the row: it keeps its own kind and puts the
//            other kind into the peer's shared memory (DSMEM) -- every spectrum entry is read from global memory once;
//   stores   the outputs are exchanged through DSMEM again so that rank c writes the contiguous half
//            [c n, (c+1) n) reals of the row: 32 contiguous bytes per thread (the first version stored 16-byte pieces
//            with a stride of 32 from two CTAs: half-written sectors, 1.67 TB/s);
//   mask     optional bit per (row, block of 64 reals): blocks whose bit is clear are not stored, rows without any
//            bit are skipped altogether (the NUFFT fine grid is only read around the interface points).
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_c2r_rows_long(const cplx* __restrict__ tw, const cplx* __restrict__ twn, const cplx* __restrict__ half_tw,
                const cplx* __restrict__ in, double* __restrict__ out, int64_t in_pitch, int64_t out_pitch, double scale,
                int nz, int64_t out_nz, const uint32_t* __restrict__ mask, int64_t mask_pitch) {
    // out_nz: only the first out_nz complex pairs of an output row are needed; the others are not written
    // nz: only the first nz entries of an input row can be nonzero (zero-padded spectra: the NUFFT fine grid); the
    // others are not read
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    static_assert(G::L == 1 && G::EPT == 16 && K::R0 == 16, "pair kernel: one line per CTA, radix-16 first stage");
    constexpr int RL = G::RLAST;
    constexpr int m = N, n = 2 * N, H = N / 2, Q = N / 16;
    extern __shared__ cplx smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int jt = threadIdx.x;
    const int64_t row = blockIdx.x >> 1;
    const uint32_t* mrow = mask ? mask + row * mask_pitch : nullptr;
    if (mrow) {
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < N / 512; ++w) any |= mrow[w];
        if (!any) return;                              // both CTAs of the pair take the same decision
    }
    cplx* line = smem;
    cplx* peer = cluster.map_shared_rank(line, c ^ 1);
    cluster.sync();                                    // the peer CTA is running: its shared memory may be written
    const cplx* src = in + row * in_pitch;
    cplx keep[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = jt + (8 * c + q) * Q;
        // Z[kk] = E + i O,  E = (X[kk] + conj X[n-kk])/2,  O = (X[kk] - conj X[n-kk])/2 * conj(half_tw[kk])
        const cplx zero = make_double2(0.0, 0.0);
        const cplx x0 = (k < nz) ? src[k] : zero, c0 = (n - k < nz) ? cconj(src[n - k]) : zero;
        const cplx x1 = (k + m < nz) ? src[k + m] : zero, c1 = (m - k < nz) ? cconj(src[m - k]) : zero;
        const cplx z0 = cadd(cscale(cadd(x0, c0), 0.5), cmuli(cmulc(cscale(csub(x0, c0), 0.5), half_tw[k])));
        const cplx z1 = cadd(cscale(cadd(x1, c1), 0.5), cmuli(cmulc(cscale(csub(x1, c1), 0.5), half_tw[k + m])));
        const cplx A = cadd(z0, z1), B = cmulc(csub(z0, z1), twn[k]);
        keep[q] = c ? B : A;
        peer[jt + q * Q] = c ? A : B;
    }
    cluster.sync();
    cplx v[G::EPT];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const cplx o = line[jt + q * Q];
        v[q] = c ? o : keep[q];
        v[8 + q] = c ? keep[q] : o;
    }
    reg_fft<N, +1, true>(v, line, jt, tw);
    // y_c[e] = z[2e + c] (packed sample = two reals).  Plane c of the owner of e's half <- y_c
    cluster.sync();
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) {
        const int e = reg_elem<N, RL>(jt, s);
        const bool upper = (s % RL) >= RL / 2;         // e >= H, known at compile time
        cplx* dst = ((int)upper == c) ? line : peer;
        dst[c * H + (e - (upper ? H : 0))] = cscale(v[s], scale);
    }
    cluster.sync();
    double* orow = out + row * out_pitch;
#pragma unroll
    for (int q = 0; q < H / G::NT; ++q) {
        const int e1 = jt + q * G::NT;
        const int e = c * H + e1;                      // packed samples 2e, 2e+1 = reals [4e, 4e+4)
        if (mrow && !((mrow[e >> 9] >> ((e >> 4) & 31)) & 1u)) continue;
        if (2 * (int64_t)e < out_nz) *reinterpret_cast<cplx*>(orow + 4 * (int64_t)e) = line[e1];
        if (2 * (int64_t)e + 1 < out_nz) *reinterpret_cast<cplx*>(orow + 4 * (int64_t)e + 2) = line[H + e1];
    }
}
template <int N>
static int launch_c2r_long(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                           double scale, int nz, int64_t out_nz, const uint32_t* mask, int64_t mask_pitch, cudaStream_t st) {
    using G = RegGeom<N>;
    int rc = set_smem(k_c2r_rows_long<N>, G::SMEM);
    if (rc) return rc;
    for (int64_t r0 = 0; r0 < batch; r0 += (1 << 29)) {
        const int64_t nb = std::min<int64_t>(batch - r0, 1 << 29);
        k_c2r_rows_long<N><<<(unsigned)(2 * nb), G::NT, G::SMEM, st>>>(p->half->d.tw, p->d.tw, p->half_tw_dev, in + r0 * in_pitch,
                                                                      out + r0 * out_pitch, in_pitch, out_pitch, scale, nz, out_nz,
                                                                      mask ? mask + r0 * mask_pitch : nullptr, mask_pitch);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}

static bool fft_split_size(int n) { return n % 2 == 0 && fft_reg_size(n / 2) && n / 2 >= 4096; }
// smallest R <= 16 with n / R a register-kernel length (3*2^k lines, power-of-two lines longer than 8192)
static int fft_split_radix(int n) {
    if (fft_reg_size(n)) return 0;
    for (int R = 2; R <= 16; ++R)
        if (n % R == 0 && fft_reg_size(n / R) && (n / R >= 1024 || n <= 8192)) return R;
    return 0;
}
static int ensure_split(FftPlan* p, int R) {
    if (p->half && p->splitR == R) return IPDE_OK;
    fft_plan_destroy(p->half);
    p->half = nullptr;
    p->splitR = R;
    return fft_plan_create(p->d.n / R, &p->half);
}
static int ensure_half(FftPlan* p) { return ensure_split(p, 2); }
static int rows_reg_sub(FftPlan* p, const cplx* in, cplx* out, int64_t nsub, int64_t in_pitch, int64_t in_sub, int in_es,
                        int64_t out_pitch, int64_t out_sub, int out_es, int sign, double scale, cudaStream_t st,
                        const double* mul = nullptr, int64_t mul_pitch = 0, int64_t in_nz = (int64_t)1 << 60) {
    const int R = p->splitR;
    const int m = p->d.n / R;
#define CALL_SUB(N_) return launch_rows_reg_sub<N_>(p->half, in, out, nsub, R, in_pitch, in_sub, in_es, out_pitch, out_sub, out_es, sign, scale, mul, mul_pitch, in_nz, st)
    REG_DISPATCH(m, CALL_SUB);
#undef CALL_SUB
}

// Radix-R steps around the m-point sub-transforms (n = R m), all indices natural:
//   combine (decimation in time, after the sub-transforms E_s = FFT_m(z[R i + s])):
//       Z[k' + m r] = sum_s (w_n^{s k'} E_s[k']) w_R^{s r}
//   split (decimation in frequency, before the sub-transforms):
//       A_s[k'] = w_n^{s k'} sum_r Z[k' + m r] w_R^{s r},   then z[R i + s] = FFT_m(A_s)[i]
// (w = exp(-/+ 2 pi i / .) for sign -/+).  One thread per k': 2R complex loads/stores, O(R^2) flops.
template <bool SPLIT>
__global__ void k_radix_step(int m, int R, const cplx* __restrict__ twn, const cplx* __restrict__ in, cplx* __restrict__ out,
                             int64_t batch, int64_t in_pitch, int64_t out_pitch, int sign, double scale) {
    const int64_t total = batch * m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / m;
        const int k = (int)(i - row * m);
        const cplx* src = in + row * in_pitch;
        cplx* dst = out + row * out_pitch;
        cplx x[16];
        for (int s = 0; s < R; ++s) {
            cplx v = src[k + (int64_t)m * s];
            if (!SPLIT && s > 0) {
                const cplx w = twn[(int64_t)s * k];                  // s k < R m = n
                v = (sign < 0) ? cmul(v, w) : cmulc(v, w);
            }
            x[s] = v;
        }
        for (int r = 0; r < R; ++r) {
            cplx acc = x[0];
            for (int s = 1; s < R; ++s) {
                const cplx w = twn[(int64_t)((s * r) % R) * m];      // w_R^{s r}
                acc = cadd(acc, (sign < 0) ? cmul(x[s], w) : cmulc(x[s], w));
            }
            if (SPLIT && r > 0) {
                const cplx w = twn[(int64_t)r * k];
                acc = (sign < 0) ? cmul(acc, w) : cmulc(acc, w);
            }
            dst[k + (int64_t)m * r] = cscale(acc, scale);
        }
    }
}
static int radix_step(bool split, FftPlan* p, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                      int sign, double scale, cudaStream_t st) {
    const int R = p->splitR, m = p->d.n / R;
    const int blocks = (int)std::min<int64_t>(cdiv(batch * m, 256), (int64_t)ipde_num_sms() * 16);
    KERNEL_TAG("k_radix_step", 32.0 * p->d.n * (double)batch);
    if (split) k_radix_step<true><<<blocks, 256, 0, st>>>(m, R, p->d.tw, in, out, batch, in_pitch, out_pitch, sign, scale);
    else k_radix_step<false><<<blocks, 256, 0, st>>>(m, R, p->d.tw, in, out, batch, in_pitch, out_pitch, sign, scale);
    KERNEL_CHECK();
    return IPDE_OK;
}
// complex rows through the split: sub-transforms with strided loads into the scratch, combine into `out`
static int fft_rows_split(FftPlan* p, int R, const cplx* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                          int sign, double scale, cudaStream_t st) {
    int rc = ensure_split(p, R);
    if (rc) return rc;
    const int n = p->d.n, m = n / R;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)1 << 24) / n));
    rc = ensure_work(p, (size_t)n * chunk);
    if (rc) return rc;
    for (int64_t r0 = 0; r0 < batch; r0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, batch - r0);
        KERNEL_TAG("k_fft_rows_long", 32.0 * n * (double)nb);
        rc = rows_reg_sub(p, in + r0 * in_pitch, p->work, R * nb, in_pitch, 1, R, n, m, 1, sign, 1.0, st);
        if (rc) return rc;
        rc = radix_step(false, p, p->work, out + r0 * out_pitch, nb, n, out_pitch, sign, scale, st);
        if (rc) return rc;
    }
    return IPDE_OK;
}
// out = FFT(mul * IFFT(in)) / n for lines of length n = R m: split step, inverse sub-transforms with strided stores
// (time order), forward sub-transforms reading time order with the multiplier applied on the loads, combine
static int fft_mul_fft_rows_split(FftPlan* p, int R, const cplx* in, cplx* out, const double* mul, int64_t batch,
                                  cudaStream_t st) {
    int rc = ensure_split(p, R);
    if (rc) return rc;
    const int n = p->d.n, m = n / R;
    rc = ensure_work(p, (size_t)2 * n * batch);
    if (rc) return rc;
    cplx* w0 = p->work;
    cplx* w1 = p->work + (size_t)n * batch;
    rc = radix_step(true, p, in, w0, batch, n, n, +1, 1.0, st);
    if (rc) return rc;
    KERNEL_TAG("k_fft_rows_long", 32.0 * n * (double)batch);
    rc = rows_reg_sub(p, w0, w1, R * batch, n, m, 1, n, 1, R, +1, 1.0 / n, st);
    if (rc) return rc;
    KERNEL_TAG("k_fft_rows_long", 40.0 * n * (double)batch);
    rc = rows_reg_sub(p, w1, w0, R * batch, n, 1, R, n, m, 1, -1, 1.0, st, mul, n);
    if (rc) return rc;
    return radix_step(false, p, w0, out, batch, n, n, -1, 1.0, st);
}

// r2c of rows of 2n reals (n = 2m complex): E_0 | E_1 sit in the first 2m slots of the output row; every thread
// owns the closed orbit {k, m-k, m+k, n-k} and rewrites it in place with the half spectrum X[0..n].
__global__ void k_r2c_untangle_split(int m, const cplx* __restrict__ twn, const cplx* __restrict__ half_tw, cplx* __restrict__ out,
                                     int64_t batch, int64_t out_pitch, const SlabScatter sc) {
    const int n = 2 * m;
    const int per = m / 2 + 1;                                    // k = 0 .. m/2
    const int64_t total = batch * per;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / per;
        const int k = (int)(i % per);
        cplx* o = out + row * out_pitch;
        const int km = (k == 0) ? 0 : m - k;                      // index of E[m - k] (periodic in m)
        const cplx e0k = o[k], e1k = o[m + k], e0m = o[km], e1m = o[m + km];
        const cplx tk = cmul(twn[k], e1k);
        // w_n^{m-k} = -conj(w_n^k)
        const cplx wm = (k == 0) ? make_double2(1.0, 0.0) : make_double2(-twn[k].x, twn[k].y);
        const cplx tm = cmul(wm, e1m);
        const cplx Zk = cadd(e0k, tk), Zkm = csub(e0k, tk);       // Z[k], Z[k + m]
        const cplx Zmk = cadd(e0m, tm), Znk = csub(e0m, tm);      // Z[m - k], Z[n - k]   (k = 0: Z[0], Z[m])
        auto X = [&](cplx za, cplx zb, int kk) {                  // half-spectrum entry kk from Z[kk], Z[n - kk]
            const cplx zc = cconj(zb);
            const cplx E = cscale(cadd(za, zc), 0.5);
            const cplx O = cmulni(cscale(csub(za, zc), 0.5));
            return cadd(E, cmul(half_tw[kk], O));
        };
        if (sc.R) {
            // forward slab transpose on the stores (the sub-transforms stay in the local scratch rows)
            if (k == 0) {
                *slab_fwd_addr(sc, row, 0) = X(Zk, Zk, 0);
                *slab_fwd_addr(sc, row, n) = X(Zk, Zk, n);
                *slab_fwd_addr(sc, row, m) = X(Zkm, Zkm, m);
            } else {
                *slab_fwd_addr(sc, row, k) = X(Zk, Znk, k);
                *slab_fwd_addr(sc, row, n - k) = X(Znk, Zk, n - k);
                *slab_fwd_addr(sc, row, m - k) = X(Zmk, Zkm, m - k);
                *slab_fwd_addr(sc, row, m + k) = X(Zkm, Zmk, m + k);
            }
        } else if (k == 0) {
            // Z[0] = Zk, Z[m] = Zkm; X[0] pairs Z[0] with itself, X[n] likewise, X[m] pairs Z[m] with Z[m]
            o[0] = X(Zk, Zk, 0);
            o[n] = X(Zk, Zk, n);
            o[m] = X(Zkm, Zkm, m);
        } else {
            const cplx x_k = X(Zk, Znk, k), x_nk = X(Znk, Zk, n - k);
            const cplx x_mk = X(Zmk, Zkm, m - k), x_km = X(Zkm, Zmk, m + k);
            o[k] = x_k; o[n - k] = x_nk;
            o[m - k] = x_mk; o[m + k] = x_km;          // 2k == m: the same two entries, the same values
        }
    }
}
// r2c of rows of 2n reals (n = 2m complex), one CLUSTER OF TWO CTAs per row -- the forward counterpart of
// k_c2r_rows_long, replacing the strided sub-transforms + in-place untangle pass:
//   loads    rank c reads the contiguous half [c m, (c+1) m) of the packed row z[j] = (x[2j], x[2j+1]) and deals it out:
//            z[2i + r] goes to slot i of rank r's shared memory (DSMEM) -- every sector is read once, by one CTA;
//   FFT      rank r transforms its sub-line E_r = FFT_m(z[2i + r]) in registers;
//   stores   both E_r are left in shared memory; rank c combines (Z[k] = E_0 + w^k E_1, Z[k+m] = E_0 - w^k E_1) and
//            untangles the orbits {k, m-k, m+k, n-k} with k in [c m/4, (c+1) m/4) (+ m/4 itself and the three real
//            entries 0, m, n), reading the peer's E over DSMEM, and writes four contiguous runs of the half spectrum.
// Optional forward slab transpose on the stores (SlabScatter).
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RegGeom<N>::NT, RegGeom<N>::MINB)
k_r2c_rows_pair(const cplx* __restrict__ tw, const cplx* __restrict__ twn, const cplx* __restrict__ half_tw,
                const cplx* __restrict__ in, cplx* __restrict__ out, int64_t in_pitch, int64_t out_pitch, int64_t in_nz,
                const SlabScatter sc) {
    // in: packed rows (in_pitch in complex units); in_nz: packed entries j >= in_nz are zero and are not read
    using G = RegGeom<N>;
    using K = RegCfg<N>;
    static_assert(G::L == 1 && G::EPT == 16, "pair kernel: one line per CTA");
    constexpr int RL = G::RLAST;
    constexpr int m = N, n = 2 * N;
    extern __shared__ cplx smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int jt = threadIdx.x;
    const int64_t row = blockIdx.x >> 1;
    cplx* line = smem;
    cplx* peer = cluster.map_shared_rank(line, c ^ 1);
    cluster.sync();                                    // the peer CTA is running: its shared memory may be written
    const cplx* src = in + row * in_pitch;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int j = c * m + jt + q * G::NT;          // packed index, contiguous across the CTA
        const cplx z = (j < in_nz) ? src[j] : make_double2(0.0, 0.0);
        cplx* dst = ((j & 1) == c) ? line : peer;
        dst[j >> 1] = z;
    }
    cluster.sync();
    cplx v[G::EPT];
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) v[s] = line[reg_elem<N, K::R0>(jt, s)];
    reg_fft<N, -1, true>(v, line, jt, tw);
    __syncthreads();                                   // the last exchange has been read
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) line[reg_elem<N, RL>(jt, s)] = v[s];
    cluster.sync();
    const cplx* E0 = c ? peer : line;
    const cplx* E1 = c ? line : peer;
    cplx* o = out + row * out_pitch;
    auto X = [&](cplx za, cplx zb, int kk) {           // half-spectrum entry kk from Z[kk], Z[n - kk]
        const cplx zc = cconj(zb);
        const cplx E = cscale(cadd(za, zc), 0.5);
        const cplx O = cmulni(cscale(csub(za, zc), 0.5));
        return cadd(E, cmul(half_tw[kk], O));
    };
    auto put = [&](int kk, cplx val) {
        if (sc.R) *slab_fwd_addr(sc, row, kk) = val; else o[kk] = val;
    };
    // orbits k in [c m/4, (c+1) m/4); rank 1 also takes k = m/2 (the orbit {m/2, 3m/2})
    for (int t = jt; t < m / 4 + c; t += G::NT) {
        const int k = c * (m / 4) + t;
        const int km = (k == 0) ? 0 : m - k;
        const cplx e0k = E0[k], e1k = E1[k], e0m = E0[km], e1m = E1[km];
        const cplx tk = cmul(twn[k], e1k);
        const cplx wm = (k == 0) ? make_double2(1.0, 0.0) : make_double2(-twn[k].x, twn[k].y);     // w_n^{m-k}
        const cplx tm = cmul(wm, e1m);
        const cplx Zk = cadd(e0k, tk), Zkm = csub(e0k, tk);       // Z[k], Z[k + m]
        const cplx Zmk = cadd(e0m, tm), Znk = csub(e0m, tm);      // Z[m - k], Z[n - k]   (k = 0: Z[0], Z[m])
        if (k == 0) {
            put(0, X(Zk, Zk, 0));
            put(n, X(Zk, Zk, n));
            put(m, X(Zkm, Zkm, m));
        } else {
            put(k, X(Zk, Znk, k));
            put(n - k, X(Znk, Zk, n - k));
            put(m - k, X(Zmk, Zkm, m - k));
            put(m + k, X(Zkm, Zmk, m + k));                      // 2k == m: the same two entries, the same values
        }
    }
    cluster.sync();                                    // the peer may still be reading this CTA's shared memory
}
template <int N>
static int launch_r2c_pair(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                           int64_t in_nz, const SlabScatter& sc, cudaStream_t st) {
    using G = RegGeom<N>;
    int rc = set_smem(k_r2c_rows_pair<N>, G::SMEM);
    if (rc) return rc;
    for (int64_t r0 = 0; r0 < batch; r0 += (1 << 29)) {
        const int64_t nb = std::min<int64_t>(batch - r0, 1 << 29);
        SlabScatter s2 = sc;
        s2.row0 += r0;
        k_r2c_rows_pair<N><<<(unsigned)(2 * nb), G::NT, G::SMEM, st>>>(p->half->d.tw, p->d.tw, p->half_tw_dev,
                                                                      reinterpret_cast<const cplx*>(in) + r0 * (in_pitch / 2),
                                                                      out + r0 * out_pitch, in_pitch / 2, out_pitch, in_nz, s2);
        KERNEL_CHECK();
    }
    return IPDE_OK;
}
// Measured (scripts/bench_rows.py, 16384^2 solve): the fused pair kernel is 5-14 % SLOWER than sub-transforms + untangle
// pass (r2c of 16384 rows x 16384 reals 2.26 vs 1.98 ms; 5.8 vs 5.2 ms inside the solve) -- the 8192-point CTA is bound by
// its own latency (one CTA per SM), and the two extra trips through shared memory cost more than the saved pass.  Kept
// as an option (IPDE_FFT_PAIR=1); the default is the two-kernel path.
static const bool g_fft_pair = []() { const char* e = getenv("IPDE_FFT_PAIR"); return e && e[0] == '1'; }();

// Global-memory versions of the pack / unpack steps for long rows
__global__ void k_r2c_untangle(int n, const cplx* __restrict__ half_tw, const cplx* __restrict__ z, cplx* __restrict__ out,
                               int64_t batch, int64_t z_pitch, int64_t out_pitch) {
    const int64_t total = batch * (n + 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (n + 1);
        const int k = (int)(i % (n + 1));
        const cplx* zz = z + row * z_pitch;
        cplx zk = zz[k == n ? 0 : k];
        cplx zc = cconj(zz[k == 0 ? 0 : n - k]);
        cplx E = cscale(cadd(zk, zc), 0.5);
        cplx O = cmulni(cscale(csub(zk, zc), 0.5));
        out[row * out_pitch + k] = cadd(E, cmul(half_tw[k], O));
    }
}
__global__ void k_c2r_tangle(int n, const cplx* __restrict__ half_tw, const cplx* __restrict__ in, cplx* __restrict__ z,
                             int64_t batch, int64_t in_pitch, int64_t z_pitch) {
    const int64_t total = batch * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / n;
        const int k = (int)(i % n);
        const cplx* src = in + row * in_pitch;
        cplx xk = src[k];
        cplx xc = cconj(src[n - k]);
        cplx E = cscale(cadd(xk, xc), 0.5);
        cplx O = cmulc(cscale(csub(xk, xc), 0.5), half_tw[k]);
        z[row * z_pitch + k] = cadd(E, cmuli(O));
    }
}

int fft_r2c_rows(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 cudaStream_t st) {
    return fft_r2c_rows_nz(p, in, out, batch, in_pitch, out_pitch, 2 * (int64_t)p->d.n, st);
}

int fft_r2c_rows_nz(FftPlan* p, const double* in, cplx* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                    int64_t in_nz, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (!p->fast && !p->any && n <= IPDE_FFT_GENERIC_MAX) return fft_r2c_rows_generic(p, in, out, batch, in_pitch, out_pitch, st);
    if (g_fft_reg && fft_split_size(n)) {
        int rc = ensure_half(p);
        if (rc) return rc;
        const int m = n / 2;
        KERNEL_TAG("k_r2c_rows_long", (8.0 * std::min<int64_t>(2 * (int64_t)n, in_nz) + 16.0 * (n + 1)) * (double)batch);
        if (g_fft_pair && in_pitch % 2 == 0) {
            if (m == 4096) return launch_r2c_pair<4096>(p, in, out, batch, in_pitch, out_pitch, (in_nz + 1) / 2, slab_scatter_none(), st);
            return launch_r2c_pair<8192>(p, in, out, batch, in_pitch, out_pitch, (in_nz + 1) / 2, slab_scatter_none(), st);
        }
        // packed row z[j] = (x[2j], x[2j+1]); sub-line r = z[2i + r]: element stride 2, offset r; reals beyond in_nz
        // are zero (zero-padded rows of the evaluator box) and are not read
        rc = rows_reg_sub(p, reinterpret_cast<const cplx*>(in), out, 2 * batch, in_pitch / 2, 1, 2, out_pitch, m, 1, -1, 1.0, st,
                          nullptr, 0, (in_nz + 1) / 2);
        if (rc) return rc;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * (m / 2 + 1), 256), (int64_t)ipde_num_sms() * 16);
        KERNEL_TAG("k_r2c_untangle_split", 32.0 * n * (double)batch);
        k_r2c_untangle_split<<<blocks, 256, 0, st>>>(m, p->d.tw, p->half_tw_dev, out, batch, out_pitch, slab_scatter_none());
        KERNEL_CHECK();
        return IPDE_OK;
    }
    if (n > IPDE_FFT_SMEM_MAX || !p->fast) {
        // complex FFT of the packed rows into the front of each output row, then untangle in place is not
        // possible (aliasing), so go through the plan's work buffer
        int rc = ensure_work(p, (size_t)3 * n * batch);
        if (rc) return rc;
        cplx* z = p->work + (size_t)2 * n * batch;
        rc = fft_lines_global(p, reinterpret_cast<const cplx*>(in), z, batch, in_pitch / 2, 1, n, 1, -1, 1.0, st);
        if (rc) return rc;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * (n + 1), 256), (int64_t)ipde_num_sms() * 16);
        k_r2c_untangle<<<blocks, 256, 0, st>>>(n, p->half_tw_dev, z, out, batch, n, out_pitch);
        KERNEL_CHECK();
        return IPDE_OK;
    }
    KERNEL_TAG("k_r2c_rows", (16.0 * n + 16.0 * (n + 1)) * (double)batch);
    if (g_fft_reg && fft_reg_size(n)) {
#define CALL_REG(N_) return launch_r2c_reg<N_>(p, in, out, batch, in_pitch, out_pitch, st)
        REG_DISPATCH(n, CALL_REG);
#undef CALL_REG
    }
    const int C = rows_per_tile(n, batch);
    const size_t smem = fft_tile_smem((int64_t)n * C);
    const int ept = tile_ept((int64_t)n * C);
    const unsigned g = (unsigned)cdiv(batch, C);
    const int t = fft_tile_threads(n, C, ept);
    int rc = IPDE_OK;
    FFT_TILE_LAUNCH(k_r2c_rows, ept, p->pow2, g, t, smem, st, p->d, p->half_tw_dev, in, out, batch, in_pitch, out_pitch, C);
    if (rc) return rc;
    KERNEL_CHECK();
    return IPDE_OK;
}

int fft_r2c_rows_scatter(FftPlan* p, const double* in, cplx* scratch, int64_t batch, int64_t in_pitch, int64_t scratch_pitch,
                         const SlabScatter& sc, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    ARG_CHECK(sc.R >= 1 && sc.R <= IPDE_SLAB_MAX_RANKS && sc.Pc * sc.R >= n + 1, "r2c scatter: bad slab description");
    if (g_fft_reg && p->fast && fft_split_size(n)) {
        int rc = ensure_half(p);
        if (rc) return rc;
        const int m = n / 2;
        KERNEL_TAG("k_r2c_rows_long", (16.0 * n + 16.0 * n) * (double)batch);
        if (g_fft_pair && in_pitch % 2 == 0) {
            if (m == 4096) return launch_r2c_pair<4096>(p, in, scratch, batch, in_pitch, scratch_pitch, n, sc, st);
            return launch_r2c_pair<8192>(p, in, scratch, batch, in_pitch, scratch_pitch, n, sc, st);
        }
        rc = rows_reg_sub(p, reinterpret_cast<const cplx*>(in), scratch, 2 * batch, in_pitch / 2, 1, 2, scratch_pitch, m, 1, -1, 1.0,
                          st, nullptr, 0, n);
        if (rc) return rc;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * (m / 2 + 1), 256), (int64_t)ipde_num_sms() * 16);
        KERNEL_TAG("k_r2c_untangle_split", 32.0 * n * (double)batch);
        k_r2c_untangle_split<<<blocks, 256, 0, st>>>(m, p->d.tw, p->half_tw_dev, scratch, batch, scratch_pitch, sc);
        KERNEL_CHECK();
        return IPDE_OK;
    }
    if (g_fft_reg && p->fast && fft_reg_size(n) && n <= IPDE_FFT_SMEM_MAX) {
        KERNEL_TAG("k_r2c_rows", (16.0 * n + 16.0 * (n + 1)) * (double)batch);
#define CALL_REG(N_) return launch_r2c_reg<N_>(p, in, scratch, batch, in_pitch, scratch_pitch, st, sc)
        REG_DISPATCH(n, CALL_REG);
#undef CALL_REG
    }
    ipde_set_error("r2c scatter: rows of %d reals have no register-resident kernel", 2 * n);
    return IPDE_ERR_UNSUPPORTED;
}

int fft_c2r_rows(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                 double scale, cudaStream_t st) {
    return fft_c2r_rows_nz(p, in, out, batch, in_pitch, out_pitch, scale, p->d.n + 1, st, 2 * (int64_t)p->d.n);
}

int fft_c2r_rows_nz(FftPlan* p, const cplx* in, double* out, int64_t batch, int64_t in_pitch, int64_t out_pitch,
                    double scale, int nz, cudaStream_t st, int64_t out_nz, const uint32_t* mask, int64_t mask_pitch) {
    if (out_nz <= 0) out_nz = 2 * (int64_t)p->d.n;
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (!p->fast && !p->any && n <= IPDE_FFT_GENERIC_MAX) return fft_c2r_rows_generic(p, in, out, batch, in_pitch, out_pitch, scale, st);
    if (g_fft_reg && fft_split_size(n)) {
        int rc = ensure_half(p);
        if (rc) return rc;
        const int m = n / 2;
        // masked stores: the written bytes depend on the mask and are not counted
        KERNEL_TAG("k_c2r_rows_long", ((mask ? 0.0 : 8.0 * std::min<int64_t>(2 * (int64_t)n, out_nz)) + 16.0 * std::min(n + 1, nz)) * (double)batch);
        if (m == 4096) return launch_c2r_long<4096>(p, in, out, batch, in_pitch, out_pitch, scale, nz, (out_nz + 1) / 2, mask, mask_pitch, st);
        return launch_c2r_long<8192>(p, in, out, batch, in_pitch, out_pitch, scale, nz, (out_nz + 1) / 2, mask, mask_pitch, st);
    }
    if (n > IPDE_FFT_SMEM_MAX || !p->fast) {
        int rc = ensure_work(p, (size_t)3 * n * batch);
        if (rc) return rc;
        cplx* z = p->work + (size_t)2 * n * batch;
        const int blocks = (int)std::min<int64_t>(cdiv(batch * n, 256), (int64_t)ipde_num_sms() * 16);
        k_c2r_tangle<<<blocks, 256, 0, st>>>(n, p->half_tw_dev, in, z, batch, in_pitch, n);
        KERNEL_CHECK();
        return fft_lines_global(p, z, reinterpret_cast<cplx*>(out), batch, n, 1, out_pitch / 2, 1, +1, scale, st);
    }
    KERNEL_TAG("k_c2r_rows", (16.0 * n + 16.0 * (n + 1)) * (double)batch);
    if (g_fft_reg && fft_reg_size(n)) {
        if (n < 1024) mask = nullptr;                  // rows shorter than one mask word
#define CALL_REG(N_) return launch_c2r_reg<N_>(p, in, out, batch, in_pitch, out_pitch, scale, mask, mask_pitch, st)
        REG_DISPATCH(n, CALL_REG);
#undef CALL_REG
    }
    const int C = rows_per_tile(n, batch);
    const size_t smem = fft_tile_smem((int64_t)n * C);
    const int ept = tile_ept((int64_t)n * C);
    const unsigned g = (unsigned)cdiv(batch, C);
    const int t = fft_tile_threads(n, C, ept);
    int rc = IPDE_OK;
    FFT_TILE_LAUNCH(k_c2r_rows, ept, p->pow2, g, t, smem, st, p->d, p->half_tw_dev, in, out, batch, in_pitch, out_pitch, C, scale);
    if (rc) return rc;
    KERNEL_CHECK();
    return IPDE_OK;
}

// ------------------------------------------------------------------------------------------
// out = FFT(m * IFFT(in)) / n, fused in shared memory, C lines per CTA
template <int EPT, int CLS, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 2 : 1)
k_fft_mul_fft_rows(FftPlanDev P, const cplx* __restrict__ in, cplx* __restrict__ out, const double* __restrict__ m,
                   int64_t batch, int C) {
    extern __shared__ cplx smem[];
    const int n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int64_t l0 = (int64_t)blockIdx.x * C;
    const int cw = (int)min((int64_t)C, batch - l0);
    for (int c = 0; c < C; ++c) {
        const int64_t off = (l0 + c) * n;
        for (int e = tid; e < n; e += nt) smem[PADI(c * n + e)] = (c < cw) ? in[off + e] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    tile_fft<0, EPT, CLS>(smem, P, C, 0, +1, tid, nt);
    const double inv = 1.0 / n;
    for (int c = 0; c < cw; ++c) {
        const int64_t off = (l0 + c) * n;
        for (int e = tid; e < n; e += nt) {
            const int sidx = PADI(c * n + e);
            smem[sidx] = cscale(smem[sidx], m[off + e] * inv);
        }
    }
    __syncthreads();
    tile_fft<0, EPT, CLS>(smem, P, C, 0, -1, tid, nt);
    for (int c = 0; c < cw; ++c) {
        const int64_t off = (l0 + c) * n;
        for (int e = tid; e < n; e += nt) out[off + e] = smem[PADI(c * n + e)];
    }
}

__global__ void k_scale_by_real(cplx* __restrict__ z, const double* __restrict__ m, int64_t total, double s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        z[i] = cscale(z[i], m[i] * s);
}

int fft_mul_fft_rows(FftPlan* p, const cplx* in, cplx* out, const double* m, int64_t batch, cudaStream_t st) {
    if (batch <= 0) return IPDE_OK;
    const int n = p->d.n;
    if (g_fft_reg && p->fast) {
        const int R = fft_split_radix(n);
        if (R) return fft_mul_fft_rows_split(p, R, in, out, m, batch, st);
    }
    if (!p->fast && !p->any && n <= IPDE_FFT_GENERIC_MAX) return fft_mul_fft_rows_generic(p, in, out, m, batch, st);
    if (n > IPDE_FFT_SMEM_MAX || !p->fast) {
        int rc = fft_lines_global(p, in, out, batch, n, 1, n, 1, +1, 1.0, st);
        if (rc) return rc;
        const int64_t total = batch * n;
        const int blocks = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)ipde_num_sms() * 16);
        k_scale_by_real<<<blocks, 256, 0, st>>>(out, m, total, 1.0 / n);
        KERNEL_CHECK();
        return fft_lines_global(p, out, out, batch, n, 1, n, 1, -1, 1.0, st);
    }
    const int C = rows_per_tile(n, batch);
    const size_t smem = fft_tile_smem((int64_t)n * C);
    const int ept = tile_ept((int64_t)n * C);
    const unsigned g = (unsigned)cdiv(batch, C);
    const int t = fft_tile_threads(n, C, ept);
    int rc = IPDE_OK;
    KERNEL_TAG("k_fft_mul_fft_rows", 40.0 * (double)n * (double)batch);
    if (g_fft_reg && fft_reg_size(n)) {
#define CALL_REG(N_) return launch_mul_fft_reg<N_>(p, in, out, m, batch, st)
        REG_DISPATCH(n, CALL_REG);
#undef CALL_REG
    }
    FFT_TILE_LAUNCH(k_fft_mul_fft_rows, ept, p->pow2, g, t, smem, st, p->d, in, out, m, batch, C);
    if (rc) return rc;
    KERNEL_CHECK();
    return IPDE_OK;
}
