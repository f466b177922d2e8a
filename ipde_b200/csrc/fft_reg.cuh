// Register-resident Stockham transforms of contiguous power-of-two lines (256 ... 8192 complex numbers).
//
// The tile engine of fft.cuh keeps the line in shared memory and runs every pass as read-sync-compute-write-sync,
// with the global load and store as two more trips (n = 1024: 8*8*4*4 -> five writes and five reads per element).
// Here the elements live in registers across the whole transform:
//
//     global -> registers (inputs of the first stage: element j + t*(N/R0), coalesced across the lanes)
//     stage 0 butterflies | exchange | twiddle + stage 1 | exchange | twiddle + stage 2 [| exchange | stage 3]
//     registers -> global (outputs of the last stage: j + t*(N/Rlast), coalesced)
//
// with radix-16 stages (one butterfly of 16 per thread) so that 1024 = 16*8*8, 2048 = 16*16*8 and 4096 = 16*16*16
// need two exchanges only.  The exchange buffer is skewed by one slot every R0 so that the stride-R0 scatter of the
// first exchange is conflict free; all later accesses are contiguous across the lanes.
//
// Stage with radix R after Ns = product of the earlier radices (Stockham autosort):
//     butterfly j (j < N/R): inputs x[j + t N/R], k = j mod Ns, twiddle w_{Ns R}^{k t}, outputs y[(j-k) R + k + t Ns]
#pragma once
#include "fft.cuh"

template <int N> struct RegCfg;
template <> struct RegCfg<256>  { static constexpr int EPT = 8,  R0 = 8,  R1 = 8,  R2 = 4,  R3 = 1; };
template <> struct RegCfg<512>  { static constexpr int EPT = 8,  R0 = 8,  R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct RegCfg<1024> { static constexpr int EPT = 16, R0 = 16, R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct RegCfg<2048> { static constexpr int EPT = 16, R0 = 16, R1 = 16, R2 = 8,  R3 = 1; };
template <> struct RegCfg<4096> { static constexpr int EPT = 16, R0 = 16, R1 = 16, R2 = 16, R3 = 1; };
template <> struct RegCfg<8192> { static constexpr int EPT = 16, R0 = 16, R1 = 8,  R2 = 8,  R3 = 8; };

template <int N>
struct RegGeom {
    using K = RegCfg<N>;
    static constexpr int EPT = K::EPT;
    static constexpr int T = N / EPT;                         // threads per line
    static constexpr int L = (T >= 128) ? 1 : 128 / T;        // lines per CTA
    static constexpr int NT = T * L;                          // threads per CTA
    static constexpr int SH = (K::R0 == 16) ? 4 : 3;          // skew: one slot every R0
    static constexpr int LINE = N + (N >> SH) + 1;            // shared slots per line
    static constexpr int RLAST = (K::R3 > 1) ? K::R3 : K::R2;
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)LINE * L;
    static constexpr int MINB = (NT <= 128) ? 4 : (NT <= 256) ? 2 : 1;   // CTAs per SM at 128 registers per thread
    static constexpr int LOGT = (T == 32) ? 5 : (T == 64) ? 6 : (T == 128) ? 7 : (T == 256) ? 8 : 9;
};
template <int SH>
__device__ __forceinline__ int reg_pad(int u) { return u + (u >> SH); }

// element index of slot s for the inputs of a radix-R stage / the outputs of the LAST stage (same formula)
template <int N, int R>
__device__ __forceinline__ int reg_elem(int jt, int s) {
    return (jt + (s / R) * RegGeom<N>::T) + (s % R) * (N / R);
}

template <int N, int R, int NS, int SIGN>
__device__ __forceinline__ void reg_butterflies(cplx* v, int jt, const cplx* __restrict__ tw) {
    using G = RegGeom<N>;
    FftPlanDev dummy;
#pragma unroll
    for (int q = 0; q < G::EPT / R; ++q) {
        if (NS > 1) {
            const int j = jt + q * G::T;
            const int k = j & (NS - 1);
            constexpr int tstep = N / (NS * R);
            if (R <= 8) {
                cplx w[R];
                w[1] = tw[k * tstep];
                if (SIGN > 0) w[1] = cconj(w[1]);
#pragma unroll
                for (int t = 2; t < R; ++t) w[t] = cmul(w[t >> 1], w[t - (t >> 1)]);
#pragma unroll
                for (int t = 1; t < R; ++t) v[q * R + t] = cmul(v[q * R + t], w[t]);
            } else {
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    const cplx w = tw[k * t * tstep];
                    v[q * R + t] = (SIGN > 0) ? cmulc(v[q * R + t], w) : cmul(v[q * R + t], w);
                }
            }
        }
        dft_small<R>(&v[q * R], SIGN, dummy);
    }
}

// outputs of a radix-RP stage (after NS earlier points) -> shared line -> inputs of a radix-R stage
template <int N, int RP, int NS, int R, bool SYNC_FIRST>
__device__ __forceinline__ void reg_exchange(cplx* v, cplx* __restrict__ line, int jt) {
    using G = RegGeom<N>;
    if (SYNC_FIRST) __syncthreads();
#pragma unroll
    for (int q = 0; q < G::EPT / RP; ++q) {
        const int j = jt + q * G::T;
        const int k = j & (NS - 1);
        const int b = (j - k) * RP + k;
#pragma unroll
        for (int t = 0; t < RP; ++t) line[reg_pad<G::SH>(b + t * NS)] = v[q * RP + t];
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < G::EPT; ++s) v[s] = line[reg_pad<G::SH>(reg_elem<N, R>(jt, s))];
}

// v: inputs of stage 0 (slot s <-> element reg_elem<N, R0>) -> outputs of the last stage (reg_elem<N, RLAST>).
// `dirty`: the line buffer may still be read by other threads (a previous use in the same kernel).
template <int N, int SIGN, bool DIRTY>
__device__ __forceinline__ void reg_fft(cplx* v, cplx* __restrict__ line, int jt, const cplx* __restrict__ tw) {
    using K = RegCfg<N>;
    reg_butterflies<N, K::R0, 1, SIGN>(v, jt, tw);
    reg_exchange<N, K::R0, 1, K::R1, DIRTY>(v, line, jt);
    reg_butterflies<N, K::R1, K::R0, SIGN>(v, jt, tw);
    reg_exchange<N, K::R1, K::R0, K::R2, true>(v, line, jt);
    reg_butterflies<N, K::R2, K::R0 * K::R1, SIGN>(v, jt, tw);
    if constexpr (K::R3 > 1) {
        reg_exchange<N, K::R2, K::R0 * K::R1, K::R3, true>(v, line, jt);
        reg_butterflies<N, K::R3, K::R0 * K::R1 * K::R2, SIGN>(v, jt, tw);
    }
}

static inline bool fft_reg_size(int n) {
    return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096 || n == 8192;
}
bool fft_reg_enabled();

#define REG_DISPATCH(n, CALL)                 \
    do {                                      \
        if ((n) == 256) { CALL(256); }        \
        else if ((n) == 512) { CALL(512); }   \
        else if ((n) == 1024) { CALL(1024); } \
        else if ((n) == 2048) { CALL(2048); } \
        else if ((n) == 4096) { CALL(4096); } \
        else { CALL(8192); }                  \
    } while (0)
