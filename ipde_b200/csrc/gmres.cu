#include "gmres.cuh"
#include <algorithm>
#include <complex>

typedef std::complex<double> hc;

static const int DOT_THREADS = 256;

__device__ __forceinline__ cplx block_reduce_cplx(cplx v, cplx* sh) {
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    cplx r = make_double2(0.0, 0.0);
    if (threadIdx.x < 32) {
        r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : make_double2(0.0, 0.0);
        for (int o = 16; o > 0; o >>= 1) {
            r.x += __shfl_xor_sync(0xffffffffu, r.x, o);
            r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
        }
    }
    __syncthreads();
    return r;
}

// h[i] = sum_e conj(V_i[e]) * w[e] for i < nvec, in ONE launch: every block writes its partial sums, the last block
// to finish (atomic ticket) adds the partials in block order -- deterministic, no second kernel.
__global__ void k_multi_dot(const cplx* __restrict__ V, int nvec, const cplx* __restrict__ w, int64_t len,
                            cplx* __restrict__ partial, cplx* __restrict__ h, unsigned* __restrict__ ticket) {
    __shared__ cplx sh[32];
    __shared__ bool last;
    for (int i = 0; i < nvec; ++i) {
        const cplx* v = V + (int64_t)i * len;
        cplx acc = make_double2(0.0, 0.0);
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x) {
            const cplx a = v[e], b = w[e];
            acc.x += a.x * b.x + a.y * b.y;
            acc.y += a.x * b.y - a.y * b.x;
        }
        cplx r = block_reduce_cplx(acc, sh);
        if (threadIdx.x == 0) partial[(int64_t)i * gridDim.x + blockIdx.x] = r;
    }
    __threadfence();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = gridDim.x;
    for (int i = threadIdx.x >> 5; i < nvec; i += blockDim.x >> 5) {      // one warp per vector
        const int lane = threadIdx.x & 31;
        cplx acc = make_double2(0.0, 0.0);
        for (int b = lane; b < nb; b += 32) acc = cadd(acc, partial[(int64_t)i * nb + b]);
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) h[i] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// Fused Gram-Schmidt step, one launch:
//     w -= sum_{i<nsub} hsub[i] V_i          (skipped when nsub == 0)
//     h[i] = <D_i, w>,  i < ndot              D = V, or w itself when `self` (the norm)
// Every thread owns the elements e = gtid + r*stride: it finishes its part of w before it starts on the dot products,
// so no grid-wide dependency arises.  Dot products go eight vectors at a time (one pass over w per chunk, all loads
// independent), partials per block, the last block to arrive (ticket) adds them in block order: deterministic.
__global__ void __launch_bounds__(256)
k_sub_dot(cplx* __restrict__ w, const cplx* __restrict__ V, int nsub, const cplx* __restrict__ hsub, int ndot, int self,
          int64_t len, cplx* __restrict__ partial, cplx* __restrict__ h, unsigned* __restrict__ ticket) {
    __shared__ cplx sh[8][8];
    __shared__ bool last;
    const int64_t g0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nsub > 0) {
        for (int64_t e = g0; e < len; e += stride) {
            cplx acc = w[e];
            int i = 0;
            for (; i + 4 <= nsub; i += 4) {
                const cplx v0 = V[(int64_t)i * len + e], v1 = V[(int64_t)(i + 1) * len + e];
                const cplx v2 = V[(int64_t)(i + 2) * len + e], v3 = V[(int64_t)(i + 3) * len + e];
                acc = csub(acc, cmul(hsub[i], v0));
                acc = csub(acc, cmul(hsub[i + 1], v1));
                acc = csub(acc, cmul(hsub[i + 2], v2));
                acc = csub(acc, cmul(hsub[i + 3], v3));
            }
            for (; i < nsub; ++i) acc = csub(acc, cmul(hsub[i], V[(int64_t)i * len + e]));
            w[e] = acc;
        }
    }
    for (int c0 = 0; c0 < ndot; c0 += 8) {
        const int nc = min(8, ndot - c0);
        cplx acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = make_double2(0.0, 0.0);
        for (int64_t e = g0; e < len; e += stride) {
            const cplx b = w[e];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < nc) {
                    const cplx a = self ? b : V[(int64_t)(c0 + i) * len + e];
                    acc[i].x += a.x * b.x + a.y * b.y;
                    acc[i].y += a.x * b.y - a.y * b.x;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nc) {
                for (int o = 16; o > 0; o >>= 1) {
                    acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o);
                    acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
                }
                if (lane == 0) sh[warp][i] = acc[i];
            }
        }
        __syncthreads();
        if (threadIdx.x < nc) {
            cplx r = make_double2(0.0, 0.0);
            for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) r = cadd(r, sh[wv][threadIdx.x]);
            partial[(int64_t)(c0 + threadIdx.x) * gridDim.x + blockIdx.x] = r;
        }
        __syncthreads();
    }
    __threadfence();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = gridDim.x;
    for (int i = warp; i < ndot; i += blockDim.x >> 5) {      // one warp per vector
        cplx acc = make_double2(0.0, 0.0);
        for (int b = lane; b < nb; b += 32) acc = cadd(acc, partial[(int64_t)i * nb + b]);
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) h[i] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

__global__ void k_sub_combination(cplx* __restrict__ w, const cplx* __restrict__ V, int nvec, const cplx* __restrict__ h,
                                  int64_t len) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x) {
        cplx acc = w[e];
        for (int i = 0; i < nvec; ++i) acc = csub(acc, cmul(h[i], V[(int64_t)i * len + e]));
        w[e] = acc;
    }
}

__global__ void k_add_combination(cplx* __restrict__ y, const cplx* __restrict__ V, int nvec, const cplx* __restrict__ c,
                                  int64_t len) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x) {
        cplx acc = y[e];
        for (int i = 0; i < nvec; ++i) acc = cadd(acc, cmul(c[i], V[(int64_t)i * len + e]));
        y[e] = acc;
    }
}

// out = in / sqrt(Re normsq[0])
__global__ void k_normalize(const cplx* __restrict__ in, cplx* __restrict__ out, const cplx* __restrict__ normsq, int64_t len) {
    const double s = rsqrt(normsq[0].x);
    const double inv = (normsq[0].x > 0.0) ? s : 0.0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = cscale(in[e], inv);
}

__global__ void k_residual(const cplx* __restrict__ b, const cplx* __restrict__ ax, cplx* __restrict__ r, int64_t len) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x)
        r[e] = csub(b[e], ax[e]);
}

int gmres_workspace_ensure(GmresWorkspace* ws, int64_t len, int restart) {
    if (ws->len >= len && ws->restart >= restart) return IPDE_OK;
    gmres_workspace_free(ws);
    ws->nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(len, DOT_THREADS), (int64_t)ipde_num_sms() * 4));
    CUDA_TRY(cudaMalloc(&ws->V, sizeof(cplx) * (size_t)(restart + 1) * len));
    CUDA_TRY(cudaMalloc(&ws->w, sizeof(cplx) * (size_t)len));
    CUDA_TRY(cudaMalloc(&ws->z, sizeof(cplx) * (size_t)len));
    CUDA_TRY(cudaMalloc(&ws->y, sizeof(cplx) * (size_t)len));
    CUDA_TRY(cudaMalloc(&ws->hbuf, sizeof(cplx) * (size_t)(3 * (restart + 2))));
    CUDA_TRY(cudaMalloc(&ws->partial, sizeof(cplx) * (size_t)(restart + 2) * ws->nblocks));
    CUDA_TRY(cudaMallocHost(&ws->host, sizeof(cplx) * (size_t)(3 * 3 * (restart + 2))));
    CUDA_TRY(cudaEventCreateWithFlags(&ws->ev[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ws->ev[1], cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&ws->ticket, sizeof(unsigned)));
    CUDA_TRY(cudaMemset(ws->ticket, 0, sizeof(unsigned)));
    ws->len = len;
    ws->restart = restart;
    return IPDE_OK;
}

void gmres_workspace_free(GmresWorkspace* ws) {
    cudaFree(ws->V); cudaFree(ws->w); cudaFree(ws->z); cudaFree(ws->y); cudaFree(ws->hbuf); cudaFree(ws->partial);
    cudaFree(ws->ticket);
    if (ws->host) cudaFreeHost(ws->host);
    if (ws->ev[0]) cudaEventDestroy(ws->ev[0]);
    if (ws->ev[1]) cudaEventDestroy(ws->ev[1]);
    *ws = GmresWorkspace();
}

static int dots(GmresWorkspace* ws, const cplx* V, int nvec, const cplx* w, int64_t len, cplx* h, cudaStream_t st) {
    k_multi_dot<<<ws->nblocks, DOT_THREADS, 0, st>>>(V, nvec, w, len, reinterpret_cast<cplx*>(ws->partial), h, ws->ticket);
    KERNEL_CHECK();
    return IPDE_OK;
}

static inline int ew_blocks(int64_t len) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(len, 256), (int64_t)ipde_num_sms() * 8));
}

int gmres_solve(LinearOperator* op, GmresWorkspace* ws, const cplx* b, cplx* x, int64_t len, double tol, int restart,
                int maxiter, int* iters, double* relres, cudaStream_t st) {
    if (restart < 1) restart = 20;
    if (restart > len) restart = (int)len;
    if (maxiter < 1) maxiter = 1;
    int rc = gmres_workspace_ensure(ws, len, restart);
    if (rc) return rc;
    const int R1 = ws->restart + 2;          // stride of the three small device vectors
    cplx* d_h = ws->hbuf;
    cplx* d_h2 = ws->hbuf + R1;
    cplx* d_n = ws->hbuf + 2 * R1;
    cplx* hst = ws->host;                            // two slots for the in-flight iterations ...
    cplx* hmisc = ws->host + (size_t)2 * 3 * R1;     // ... and one for norms / the small solution vector
    cplx* V = ws->V;
    const int eb = ew_blocks(len);
    int total_its = 0;
    *relres = 0.0;
    *iters = 0;

    // ||b||
    rc = dots(ws, b, 1, b, len, d_n, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(hmisc, d_n, sizeof(cplx), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const double bnorm = sqrt(hmisc[0].x);
    CUDA_TRY(cudaMemsetAsync(ws->y, 0, sizeof(cplx) * (size_t)len, st));
    if (bnorm == 0.0) {
        CUDA_TRY(cudaMemsetAsync(x, 0, sizeof(cplx) * (size_t)len, st));
        return IPDE_OK;
    }

    std::vector<hc> H((size_t)(restart + 1) * restart), cs(restart), sn(restart), g(restart + 1);
    bool converged = false;
    for (int cycle = 0; cycle < maxiter && !converged; ++cycle) {
        // r = b - A M y   (first cycle: y = 0)
        if (cycle == 0) {
            CUDA_TRY(cudaMemcpyAsync(ws->w, b, sizeof(cplx) * (size_t)len, cudaMemcpyDeviceToDevice, st));
        } else {
            rc = op->precond(ws->y, ws->z, st); if (rc) return rc;
            rc = op->apply(ws->z, ws->w, st); if (rc) return rc;
            k_residual<<<eb, 256, 0, st>>>(b, ws->w, ws->w, len);
            KERNEL_CHECK();
        }
        double beta;
        if (cycle == 0) {
            // r = b: its norm is already on the host (and still in d_n on the device)
            k_normalize<<<eb, 256, 0, st>>>(ws->w, V, d_n, len);
            KERNEL_CHECK();
            beta = bnorm;
        } else {
            rc = dots(ws, ws->w, 1, ws->w, len, d_n, st); if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(hmisc, d_n, sizeof(cplx), cudaMemcpyDeviceToHost, st));
            k_normalize<<<eb, 256, 0, st>>>(ws->w, V, d_n, len);
            KERNEL_CHECK();
            CUDA_TRY(cudaStreamSynchronize(st));
            beta = sqrt(hmisc[0].x);
        }
        if (beta <= tol * bnorm) { *relres = beta / bnorm; converged = true; break; }
        std::fill(H.begin(), H.end(), hc(0.0));
        std::fill(g.begin(), g.end(), hc(0.0));
        g[0] = beta;
        int kdone = 0;
        // Software pipeline of depth one: iteration j+1 is enqueued before the host looks at the Hessenberg column
        // of iteration j, so the device never waits for the host round trip (one iteration of speculative work is
        // discarded when convergence is detected).
        auto enqueue = [&](int j) -> int {
            cplx* vj = V + (int64_t)j * len;
            int r = op->precond(vj, ws->z, st); if (r) return r;
            r = op->apply(ws->z, ws->w, st); if (r) return r;
            // classical Gram-Schmidt, twice
            cplx* part = reinterpret_cast<cplx*>(ws->partial);
            k_sub_dot<<<ws->nblocks, DOT_THREADS, 0, st>>>(ws->w, V, 0, nullptr, j + 1, 0, len, part, d_h, ws->ticket);
            KERNEL_CHECK();
            k_sub_dot<<<ws->nblocks, DOT_THREADS, 0, st>>>(ws->w, V, j + 1, d_h, j + 1, 0, len, part, d_h2, ws->ticket);
            KERNEL_CHECK();
            k_sub_dot<<<ws->nblocks, DOT_THREADS, 0, st>>>(ws->w, V, j + 1, d_h2, 1, 1, len, part, d_n, ws->ticket);
            KERNEL_CHECK();
            k_normalize<<<eb, 256, 0, st>>>(ws->w, V + (int64_t)(j + 1) * len, d_n, len);
            KERNEL_CHECK();
            cplx* slot = hst + (size_t)(j & 1) * (3 * R1);
            CUDA_TRY(cudaMemcpyAsync(slot, ws->hbuf, sizeof(cplx) * (size_t)(2 * R1 + 1), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaEventRecord(ws->ev[j & 1], st));
            return IPDE_OK;
        };
        auto process = [&](int j) -> int {
            CUDA_TRY(cudaEventSynchronize(ws->ev[j & 1]));
            const cplx* slot = hst + (size_t)(j & 1) * (3 * R1);
            const double hn = sqrt(slot[2 * R1].x);
            for (int i = 0; i <= j; ++i)
                H[(size_t)i * restart + j] = hc(slot[i].x + slot[R1 + i].x, slot[i].y + slot[R1 + i].y);
            H[(size_t)(j + 1) * restart + j] = hn;
            for (int i = 0; i < j; ++i) {
                hc a = H[(size_t)i * restart + j], bb = H[(size_t)(i + 1) * restart + j];
                H[(size_t)i * restart + j] = cs[i] * a + sn[i] * bb;
                H[(size_t)(i + 1) * restart + j] = -std::conj(sn[i]) * a + cs[i] * bb;
            }
            hc a = H[(size_t)j * restart + j], bb = H[(size_t)(j + 1) * restart + j];
            const double den = sqrt(std::norm(a) + std::norm(bb));
            if (std::abs(a) != 0.0) {
                cs[j] = std::abs(a) / den;
                sn[j] = (a / std::abs(a)) * std::conj(bb) / den;
            } else {
                cs[j] = 0.0; sn[j] = 1.0;
            }
            H[(size_t)j * restart + j] = cs[j] * a + sn[j] * bb;
            H[(size_t)(j + 1) * restart + j] = 0.0;
            g[j + 1] = -std::conj(sn[j]) * g[j];
            g[j] = cs[j] * g[j];
            kdone = j + 1;
            ++total_its;
            const double rel = std::abs(g[j + 1]) / bnorm;
            *relres = rel;
            if (rel <= tol || hn == 0.0) converged = true;
            return IPDE_OK;
        };
        // Speculation is suspended when the residual history predicts convergence in the iteration that is still in
        // flight (geometric extrapolation of the last two reductions): then the host looks first.
        double rel_hist[2] = {1.0, 1.0};      // rel after the last two processed iterations
        int enq = 0;                          // iterations enqueued so far
        while (!converged && kdone < restart) {
            const bool in_flight = enq > kdone;
            bool speculate = enq < restart;
            if (in_flight && kdone >= 2) {
                const double rho = rel_hist[1] / rel_hist[0];
                if (rel_hist[1] * rho <= 4.0 * tol) speculate = false;
            }
            if (in_flight && enq - kdone >= 2) speculate = false;          // at most one iteration ahead
            if (speculate) { rc = enqueue(enq); if (rc) return rc; ++enq; }
            if (enq > kdone && (!speculate || enq - kdone >= 2 || enq == restart)) {
                rc = process(kdone); if (rc) return rc;
                rel_hist[0] = rel_hist[1]; rel_hist[1] = *relres;
            }
        }
        // back substitution for the small triangular system, then y += V^T yk
        std::vector<hc> yk(kdone);
        for (int i = kdone - 1; i >= 0; --i) {
            hc s = g[i];
            for (int l = i + 1; l < kdone; ++l) s -= H[(size_t)i * restart + l] * yk[l];
            yk[i] = s / H[(size_t)i * restart + i];
        }
        for (int i = 0; i < kdone; ++i) hmisc[i] = make_double2(yk[i].real(), yk[i].imag());
        CUDA_TRY(cudaMemcpyAsync(d_h, hmisc, sizeof(cplx) * (size_t)kdone, cudaMemcpyHostToDevice, st));
        k_add_combination<<<eb, 256, 0, st>>>(ws->y, V, kdone, d_h, len);
        KERNEL_CHECK();
        if (!converged) CUDA_TRY(cudaStreamSynchronize(st));   // the host staging is reused by the next cycle
    }
    *iters = total_its;
    return op->precond(ws->y, x, st);
}
