// Blocked Householder QR of a tall matrix with the trailing update on the FP64 tensor cores.
//
// Replaces LAPACK geqrf / ormqr behind TensorKit `left_orth` / `right_orth`
// (/root/reference/src/schemes/atrg3d.jl:53-56) and is the first stage of the truncated SVD the
// north star names (`svd_trunc` behind src/utility/projectors.jl:213-219): A = Q R by Householder
// reflectors, then the one-sided Jacobi SVD of the small R in shared memory (jacobi.cu), then the
// top-chi selection on the device.
//
//   * panel (NB = 32 columns): one fused kernel per column applies reflector j to the rest of the
//     panel AND accumulates the dot products reflector j+1 needs, so a column costs one pass over
//     the panel (which stays in the 126 MB L2) and one launch; partial sums are reduced in a
//     fixed order (deterministic);
//   * compact WY: T (NB x NB) from the Gram matrix V^T V (one split-K DMMA GEMM) by a forward
//     recurrence (larft);
//   * trailing update A2 <- (I - V T^T V^T) A2 as three DMMA GEMMs (gemm_dmma.cu / gemm_tma.cu);
//   * Q is applied (never formed alone) through the stored reflectors: Y <- Q Y, two GEMMs per
//     panel.
#include "tensor.hpp"

namespace tnr {
namespace {

constexpr int NB = 32;
constexpr int QT = 256;   // threads per CTA of the panel kernels

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Grid reduction of NB per-thread accumulators: every CTA writes its partial sums, the LAST CTA
// to arrive adds all partials in a fixed order (deterministic, whichever CTA that is) and writes
// dots[t] for t in [t0, pw).
__device__ __forceinline__ void grid_reduce_nb(const double (&acc)[NB], int t0, int pw,
                                               double* partial, unsigned* counter, double* dots) {
    __shared__ double red[QT / 32][NB];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        if (t >= t0 && t < pw) {          // block-uniform condition
            const double s = warp_sum(acc[t]);
            if (lane == 0) red[warp][t] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < NB && (int)threadIdx.x >= t0 && (int)threadIdx.x < pw) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < QT / 32; ++w) s += red[w][threadIdx.x];
        partial[(long long)blockIdx.x * NB + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    {
        const int t = lane, part = warp;  // 8 parts x 32 columns
        double s = 0.0;
        if (t >= t0 && t < pw)
            for (int b = part; b < (int)gridDim.x; b += QT / 32) s += partial[(long long)b * NB + t];
        red[part][t] = s;
    }
    __syncthreads();
    if (threadIdx.x < NB && (int)threadIdx.x >= t0 && (int)threadIdx.x < pw) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < QT / 32; ++w) s += red[w][threadIdx.x];
        dots[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *counter = 0;
}

// Dots of column 0 of the panel with columns 0..pw-1 over the rows BELOW the diagonal row 0:
// partial[b][t] = sum_{r >= 1, r in block b} P[r,0] P[r,t];  diag[t] = P[0,t].
__global__ void __launch_bounds__(QT) qr_panel_first_kernel(const double* __restrict__ P,
                                                            long long lda, long long mrows, int pw,
                                                            double* __restrict__ partial,
                                                            unsigned* __restrict__ counter,
                                                            double* __restrict__ dots,
                                                            double* __restrict__ diag) {
    double acc[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) acc[t] = 0.0;
    for (long long r = 1 + blockIdx.x * (long long)QT + threadIdx.x; r < mrows;
         r += (long long)gridDim.x * QT) {
        const double x = P[r];
#pragma unroll
        for (int t = 0; t < NB; ++t)
            if (t < pw) acc[t] += x * P[r + t * lda];
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < pw) diag[threadIdx.x] = P[threadIdx.x * lda];
    grid_reduce_nb(acc, 0, pw, partial, counter, dots);
}

// Column j of the panel: reduce the partial dots, build reflector j (dlarfg), apply it to the
// columns t > j of the panel, store v below the diagonal, and accumulate the dots column j+1
// needs on the UPDATED values.  Rows are local to the panel (row j is the diagonal of column j).
__global__ void __launch_bounds__(QT) qr_panel_step_kernel(double* __restrict__ P, long long lda,
                                                           long long mrows, int j, int pw,
                                                           const double* __restrict__ dots_in,
                                                           const double* __restrict__ diag_in,
                                                           double* __restrict__ partial,
                                                           unsigned* __restrict__ counter,
                                                           double* __restrict__ dots_out,
                                                           double* __restrict__ diag_out,
                                                           double* __restrict__ tau) {
    __shared__ double w[NB];
    __shared__ double sc[3];   // scale = 1/(x0 - beta), tau, beta
    if (threadIdx.x < NB) {
        const int t = threadIdx.x;
        w[t] = (t >= j && t < pw) ? dots_in[t] : 0.0;   // tail dots for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double x0 = diag_in[j], tail2 = w[j];
        double beta = x0, tj = 0.0, scale = 0.0;
        if (tail2 > 0.0) {
            beta = -copysign(sqrt(x0 * x0 + tail2), x0);
            tj = (beta - x0) / beta;
            scale = 1.0 / (x0 - beta);
        }
        sc[0] = scale; sc[1] = tj; sc[2] = beta;
    }
    __syncthreads();
    const double scale = sc[0], tj = sc[1];
    if (threadIdx.x < NB) {
        const int t = threadIdx.x;
        // w[t] = v^T P[:, t] = P[j,t] + scale * (tail dot)
        if (t > j && t < pw) w[t] = diag_in[t] + scale * w[t];
    }
    __syncthreads();
    // diagonal row j: R entries of this row (one CTA writes them)
    if (blockIdx.x == 0 && threadIdx.x < NB) {
        const int t = threadIdx.x;
        if (t == j) { P[j + (long long)j * lda] = sc[2]; tau[j] = tj; }
        else if (t > j && t < pw) P[j + (long long)t * lda] = diag_in[t] - tj * w[t];
    }
    double acc[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) acc[t] = 0.0;
    const bool more = (j + 1 < pw);
    for (long long r = j + 1 + blockIdx.x * (long long)QT + threadIdx.x; r < mrows;
         r += (long long)gridDim.x * QT) {
        const double v = P[r + (long long)j * lda] * scale;
        P[r + (long long)j * lda] = v;
        const double tv = tj * v;
        double row[NB];
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            if (t > j && t < pw) {
                row[t] = P[r + (long long)t * lda] - tv * w[t];
                P[r + (long long)t * lda] = row[t];
            }
        }
        if (more) {
            if (r == j + 1) {
                // the next diagonal row, after the update
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    if (t > j && t < pw) diag_out[t] = row[t];
            } else {
                double x = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    if (t == j + 1) x = row[t];
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    if (t > j && t < pw) acc[t] += x * row[t];
            }
        }
    }
    if (more) grid_reduce_nb(acc, j + 1, pw, partial, counter, dots_out);
}

// explicit reflectors of a panel: Vp[r,c] = 0 (r < c), 1 (r == c), P[r,c] (r > c); rows local
__global__ void qr_extract_v_kernel(const double* __restrict__ P, long long lda, long long mrows,
                                    int pw, double* __restrict__ Vp, long long ldv) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (r >= mrows || c >= pw) return;
    Vp[r + (long long)c * ldv] = (r < c) ? 0.0 : (r == c ? 1.0 : P[r + (long long)c * lda]);
}

// T of the compact WY form, forward / columnwise (larft): T[i,i] = tau_i,
// T[a,i] = -tau_i * sum_{b=a}^{i-1} T[a,b] S[b,i] (a < i), S = V^T V.  One warp; lane a owns row a.
__global__ void qr_build_t_kernel(const double* __restrict__ S, const double* __restrict__ tau,
                                  int pw, double* __restrict__ T) {
    const int a = threadIdx.x;
    double row[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) row[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (i < pw && a < pw) {
            if (a == i) row[i] = tau[i];
            else if (a < i) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < NB; ++b)
                    if (b >= a && b < i) s += row[b] * S[b + i * NB];
                row[i] = -tau[i] * s;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) T[a + i * NB] = (a < pw && i < pw) ? row[i] : 0.0;
}

// R = upper triangle of the leading n x n block of A
__global__ void qr_copy_r_kernel(const double* __restrict__ A, long long lda, int n,
                                 double* __restrict__ R, long long ldr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n || c >= n) return;
    R[r + (long long)c * ldr] = (r <= c) ? A[r + (long long)c * lda] : 0.0;
}

// Y (m x k, ldy): rows [0, n) = X (n x k, ldx), rows [n, m) = 0
__global__ void qr_embed_kernel(const double* __restrict__ X, long long ldx, int n, long long m,
                                int k, double* __restrict__ Y, long long ldy) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (r >= m || c >= k) return;
    Y[r + (long long)c * ldy] = (r < n) ? X[r + (long long)c * ldx] : 0.0;
}

}  // namespace

// In place: on return the upper triangle of A holds R; w keeps the reflectors (explicit, m x n,
// unit lower trapezoidal) and the T blocks for qr_apply_q.
void qr_factor(Context* ctx, double* A, long long m, long long n, long long lda, QRWork& w) {
    TNR_CHECK(m >= n && n >= 1, "qr_factor: expects a tall matrix");
    w.m = m; w.n = n;
    w.V = DT(ctx, {m, n});
    const int npanel = (int)((n + NB - 1) / NB);
    w.T = DT(ctx, {NB, NB, npanel});
    w.tau = DT(ctx, {n});
    const int gmax = 4 * ctx->num_sms;
    double* partial = dalloc(ctx, (size_t)gmax * NB);
    double* diag = dalloc(ctx, (size_t)4 * NB);          // [2][NB] diagonal rows, [2][NB] dots
    double* dots = diag + 2 * NB;
    unsigned* counter = nullptr;
    TNR_CUDA(cudaMallocAsync((void**)&counter, sizeof(unsigned), ctx->stream));
    TNR_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    DT S(ctx, {NB, NB});
    for (int pnl = 0; pnl < npanel; ++pnl) {
        const long long j0 = (long long)pnl * NB;
        const int pw = (int)std::min<long long>(NB, n - j0);
        const long long mrows = m - j0;
        double* P = A + j0 + j0 * lda;
        const int g = (int)std::max<long long>(1, std::min<long long>(gmax, (mrows + QT - 1) / QT));
        qr_panel_first_kernel<<<g, QT, 0, ctx->stream>>>(P, lda, mrows, pw, partial, counter, dots,
                                                         diag);
        for (int j = 0; j < pw; ++j) {
            const int in = j & 1, out = in ^ 1;
            qr_panel_step_kernel<<<g, QT, 0, ctx->stream>>>(
                P, lda, mrows, j, pw, dots + in * NB, diag + in * NB, partial, counter,
                dots + out * NB, diag + out * NB, w.tau.p + j0);
        }
        TNR_CUDA(cudaGetLastError());
        ctx->ctr.launches += pw + 1;
        // explicit reflectors (rows j0.. of column block j0..): V[j0:, j0:j0+pw]
        double* Vp = w.V.p + j0 + j0 * m;
        {
            dim3 grid((unsigned)((mrows + 255) / 256), (unsigned)pw);
            qr_extract_v_kernel<<<grid, 256, 0, ctx->stream>>>(P, lda, mrows, pw, Vp, m);
            ctx->ctr.launches++;
        }
        // T from S = V^T V
        gemm(ctx, 'T', 'N', pw, pw, (int)mrows, 1.0, Vp, m, Vp, m, 0.0, S.p, NB);
        double* Tp = w.T.p + (size_t)pnl * NB * NB;
        qr_build_t_kernel<<<1, NB, 0, ctx->stream>>>(S.p, w.tau.p + j0, pw, Tp);
        ctx->ctr.launches++;
        // trailing update on the tensor cores: A2 <- A2 - V (T^T (V^T A2))
        const long long n2 = n - j0 - pw;
        if (n2 > 0) {
            double* A2 = A + j0 + (j0 + pw) * lda;
            DT W(ctx, {NB, n2}), W2(ctx, {NB, n2});
            gemm(ctx, 'T', 'N', pw, (int)n2, (int)mrows, 1.0, Vp, m, A2, lda, 0.0, W.p, NB);
            gemm(ctx, 'T', 'N', pw, (int)n2, pw, 1.0, Tp, NB, W.p, NB, 0.0, W2.p, NB);
            gemm(ctx, 'N', 'N', (int)mrows, (int)n2, pw, -1.0, Vp, m, W2.p, NB, 1.0, A2, lda);
        }
    }
    // the part of V above the panels' diagonal blocks is never read: the GEMMs start at row j0
    dfree(ctx, partial);
    dfree(ctx, diag);
    TNR_CUDA(cudaFreeAsync(counter, ctx->stream));
    ctx->ctr.qr_factorizations++;
}

void qr_copy_r(Context* ctx, const double* A, long long lda, long long n, double* R,
               long long ldr) {
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)n);
    qr_copy_r_kernel<<<grid, 128, 0, ctx->stream>>>(A, lda, (int)n, R, ldr);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

// Y (m x k, ldy) <- Q Y with Q = H_1 ... H_n of the factorization in w
void qr_apply_q(Context* ctx, const QRWork& w, double* Y, long long ldy, long long k) {
    const long long m = w.m, n = w.n;
    const int npanel = (int)((n + NB - 1) / NB);
    DT W(ctx, {NB, k}), W2(ctx, {NB, k});
    for (int pnl = npanel - 1; pnl >= 0; --pnl) {
        const long long j0 = (long long)pnl * NB;
        const int pw = (int)std::min<long long>(NB, n - j0);
        const long long mrows = m - j0;
        const double* Vp = w.V.p + j0 + j0 * m;
        const double* Tp = w.T.p + (size_t)pnl * NB * NB;
        double* Yp = Y + j0;
        gemm(ctx, 'T', 'N', pw, (int)k, (int)mrows, 1.0, Vp, m, Yp, ldy, 0.0, W.p, NB);
        gemm(ctx, 'N', 'N', pw, (int)k, pw, 1.0, Tp, NB, W.p, NB, 0.0, W2.p, NB);
        gemm(ctx, 'N', 'N', (int)mrows, (int)k, pw, -1.0, Vp, m, W2.p, NB, 1.0, Yp, ldy);
    }
}

// Y (m x k) = Q [X; 0] for an n x k matrix X
void qr_q_times(Context* ctx, const QRWork& w, const double* X, long long ldx, long long k,
                double* Y, long long ldy) {
    dim3 grid((unsigned)((w.m + 255) / 256), (unsigned)k);
    qr_embed_kernel<<<grid, 256, 0, ctx->stream>>>(X, ldx, (int)w.n, w.m, (int)k, Y, ldy);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
    qr_apply_q(ctx, w, Y, ldy, k);
}

}  // namespace tnr
