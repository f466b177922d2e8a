// Register-resident short column transforms for the passes of the four-step split (power-of-two lengths
// N = RA*RB in {16, 32, 64, 128}).
//
// The tile engine of fft.cuh stages every tile in shared memory and runs its passes there (six shared-memory
// accesses and ~180 issued instructions per element for a 64-point line).  For the short strided lines of a
// four-step pass that is the bottleneck, not HBM.  Here a thread keeps its 8 elements in registers from the global
// load to the global store:
//
//     global -> registers (8 rows of one column per thread; a warp reads 256..512-byte row segments)
//     radix-RA butterflies in registers
//     ONE exchange through shared memory (16 B store + 16 B load per element, conflict free without padding:
//         every aligned group of 8 lanes touches 128 contiguous bytes)
//     twiddle + radix-RB butterflies in registers
//     registers -> global
//
// A CTA of 128 threads owns a tile of C = 128/T adjacent columns (T = N/8 threads per line) and walks over tiles
// persistently; the loads of the next tile are issued into a second register set before the current tile is
// transformed, so the HBM latency is hidden without relying on occupancy.  Two shared buffers alternate, which
// leaves a single __syncthreads() per exchange.
//
// Index bookkeeping (line element index e, thread (c, jt), q-th butterfly of the thread):
//     layout A<RA,RB>:  slot q*RA + t holds element (jt + q*T) + t*RB        q < 8/RA, t < RA
//     layout B<RA,RB>:  slot q*RB + t holds element (jt + q*T) + t*RA        q < 8/RB, t < RB
// short_fft<RA,RB> maps layout A (time order for the radix-RA stage) to layout B (natural frequency order).
// Layout B<RA,RB> == layout A<RB,RA>, so a forward transform, a pointwise multiplier and the inverse transform
// chain with no reshuffle: short_fft<RA,RB,-1>, multiply, short_fft<RB,RA,+1>.
#pragma once
#include "fft.cuh"

template <int RA, int RB>
struct ShortCfg {
    static constexpr int N = RA * RB;
    static constexpr int EPT = (N > 64) ? 16 : 8;   // elements per thread (N = 128: one radix-16 butterfly)
    static constexpr int T = N / EPT;            // threads per line
    static constexpr int C = 128 / T;            // columns per tile (128 threads per CTA)
    static constexpr int LOGC = (C == 8) ? 3 : (C == 16) ? 4 : (C == 32) ? 5 : 6;
    static constexpr int BA = EPT / RA;          // radix-RA butterflies per thread
    static constexpr int BB = EPT / RB;
    static constexpr int TILE = N * C;           // 1024 elements (2048 for N = 128)
    // EPT 8: a second register set prefetches the next tile and two shared buffers alternate (one barrier per
    // exchange).  EPT 16: the registers are taken by the tile itself; latency is hidden by the 4 resident CTAs,
    // one shared buffer (32 KB) with a barrier before it is rewritten.
    static constexpr bool PREFETCH = (EPT == 8);
    static constexpr int NBUF = PREFETCH ? 2 : 1;
    static_assert(N == 16 || N == 32 || N == 64 || N == 128 || N == 256, "short transforms: N in {16, ..., 256}");
    static_assert(BA >= 1 && BB >= 1, "radix larger than the elements of a thread");
};

// element index held in slot s of layout A / B
template <int RA, int RB>
__device__ __forceinline__ int short_elem_a(int jt, int s) {
    return (jt + (s / RA) * ShortCfg<RA, RB>::T) + (s % RA) * RB;
}
template <int RA, int RB>
__device__ __forceinline__ int short_elem_b(int jt, int s) {
    return (jt + (s / RB) * ShortCfg<RA, RB>::T) + (s % RB) * RA;
}

// v: layout A on entry, layout B on exit.  sm: TILE elements.  tw[j] = exp(-2 pi i j / N).
// Contains one __syncthreads(); the caller alternates between two buffers from one call to the next.
template <int RA, int RB, int SIGN>
__device__ __forceinline__ void short_fft(cplx (&v)[ShortCfg<RA, RB>::EPT], cplx* __restrict__ sm, int c, int jt,
                                          const cplx* __restrict__ tw) {
    using K = ShortCfg<RA, RB>;
    FftPlanDev dummy;   // dft_small<4|8|16> never touches the plan
    if (K::NBUF == 1) __syncthreads();   // readers of the previous exchange are done
#pragma unroll
    for (int q = 0; q < K::BA; ++q) {
        dft_small<RA>(&v[q * RA], SIGN, dummy);
        const int j = jt + q * K::T;
#pragma unroll
        for (int t = 0; t < RA; ++t) sm[((j * RA + t) << K::LOGC) + c] = v[q * RA + t];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < K::BB; ++q) {
        const int j = jt + q * K::T;
#pragma unroll
        for (int t = 0; t < RB; ++t) v[q * RB + t] = sm[((j + t * RA) << K::LOGC) + c];
        if (RB <= 8) {
            // one table load, the other powers by a product tree (<= 3 roundings)
            cplx w[RB];
            w[1] = tw[j];
            if (SIGN > 0) w[1] = cconj(w[1]);
#pragma unroll
            for (int t = 2; t < RB; ++t) w[t] = cmul(w[t >> 1], w[t - (t >> 1)]);
#pragma unroll
            for (int t = 1; t < RB; ++t) v[q * RB + t] = cmul(v[q * RB + t], w[t]);
        } else {
            // radix 16: fifteen powers would not fit the register budget next to the tile; j*t < N, read them
#pragma unroll
            for (int t = 1; t < RB; ++t) {
                const cplx w = tw[j * t];
                v[q * RB + t] = (SIGN > 0) ? cmulc(v[q * RB + t], w) : cmul(v[q * RB + t], w);
            }
        }
        dft_small<RB>(&v[q * RB], SIGN, dummy);
    }
}

// tile -> (outer line, first column)
struct ShortTile {
    int o;
    int64_t base;     // o*outer_stride + c0 + c
    bool colok;
    int row_base;
    int col;          // c0 + c
};
template <int LOGC>
__device__ __forceinline__ ShortTile short_tile(int tile, int tiles_per_outer, int c, int64_t n_inner, int64_t outer_stride,
                                                int row_o) {
    ShortTile S;
    S.o = tile / tiles_per_outer;
    S.col = ((tile - S.o * tiles_per_outer) << LOGC) + c;
    S.colok = S.col < n_inner;
    S.base = (int64_t)S.o * outer_stride + S.col;
    S.row_base = S.o * row_o;
    return S;
}

bool fft_short_enabled();
static inline bool fft_short_size(int n) { return n == 16 || n == 32 || n == 64 || n == 128; }
// the plain (unfused) column pass also has a 256-point instance (16 x 16, 8 columns per tile): first/last pass of
// the four-step split of 32768-point columns
static inline bool fft_short_size_plain(int n) { return fft_short_size(n) || n == 256; }
#define SHORT_DISPATCH_PLAIN(n, CALL)        \
    do {                                     \
        if ((n) == 256) { CALL(16, 16); }    \
        else SHORT_DISPATCH(n, CALL);        \
    } while (0)

// SHORT_DISPATCH(n, CALL): CALL is a macro taking (RA, RB)
#define SHORT_DISPATCH(n, CALL)            \
    do {                                   \
        if ((n) == 128) { CALL(16, 8); }   \
        else if ((n) == 64) { CALL(8, 8); } \
        else if ((n) == 32) { CALL(8, 4); } \
        else { CALL(4, 4); }               \
    } while (0)
