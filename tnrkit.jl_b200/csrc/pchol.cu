// Rank-revealing factor G ~= L L^T of a symmetric positive semidefinite matrix: diagonally
// pivoted Cholesky, right-looking in blocks (the Schur complement is updated once per block of
// columns by a DMMA GEMM, inside a block the pending updates are applied to the one column
// that is being produced).
//
// Why it exists: the factored ATRG_3D step (tnrkit.jl_b200/atrg3d_factored.py, rfactor="gram")
// needs the R factors of `left_orth` / `right_orth` (/root/reference/src/schemes/atrg3d.jl:53-56)
// only up to their left orthogonal gauge, i.e. ANY L with L L^T = YD^T YD.  Round 2 measured the
// full eigendecomposition of those chi^2 x chi^2 Gram matrices (n = 2304 at chi = 48: one-pair-
// per-CTA Jacobi on global memory, 2303 launches per sweep) as the launch-bound part of
// configs[3]; a Cholesky factor is n launches of a memory-bound column kernel plus n / 64 GEMMs.
//
// No row exchanges: rows stay where they are and a row that has been a pivot is marked (its
// remaining diagonal entry d[i] is set to -1), so L is not triangular -- nothing downstream
// needs that.  Every CTA of a column launch recomputes the pivot (argmax of <= a few thousand
// doubles) instead of a separate selection launch; the diagonal is double buffered because CTAs
// of one launch read all of it and write their own rows.  |L[i, j]| is clamped to sqrt(d[i]), the
// Cauchy-Schwarz bound of a PSD Schur complement, so pivots at rounding level cannot amplify
// noise; the factorization stops when the largest remaining d is <= PCHOL_REL * max_i G[i, i].
#include "tensor.hpp"

namespace tnr {

namespace {

constexpr int PC_ROWS = 32;       // rows of the new column one CTA produces (lane = row)
constexpr int PC_THREADS = 256;   // 8 warps: warp = slice of the pending columns of the block
constexpr int PC_MAXB = 128;      // largest block of columns (pending updates per column)
constexpr double PCHOL_REL = 8.0 * 2.220446049250313e-16;

__device__ __forceinline__ void argmax_take(double& v, int& i, double v2, int i2) {
    if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
}

// d[i] = max(G[i, i], 0);  *thresh = PCHOL_REL * max_i d[i];  *rank = -1.   One CTA.
__global__ void __launch_bounds__(1024) pchol_init_kernel(const double* __restrict__ G, int n,
                                                          double* __restrict__ d,
                                                          double* __restrict__ thresh,
                                                          int* __restrict__ rank) {
    __shared__ double s_m[32];
    double mx = 0.0;
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double g = G[i + (long long)i * n];
        bad |= !isfinite(g);
        const double v = g > 0.0 ? g : 0.0;
        d[i] = v;
        mx = fmax(mx, v);
    }
    bad = __syncthreads_or(bad);
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, s_m[w]);
        *thresh = PCHOL_REL * m;
        *rank = bad ? -2 : -1;                    // -2: non-finite diagonal, reported by the host
    }
}

// Column j of L.  S: Schur complement as of the start of the block [k0, ...) (n x n, ld n,
// symmetric: column p is read), L: n x n (ld n), d_in / d_out: remaining diagonal (-1 = row has
// been a pivot).
__global__ void __launch_bounds__(PC_THREADS) pchol_column_kernel(
    const double* __restrict__ S, double* __restrict__ L, int n, const double* __restrict__ d_in,
    double* __restrict__ d_out, int j, int k0, const double* __restrict__ thresh,
    int* __restrict__ rank) {
    __shared__ double s_v[PC_THREADS / 32];
    __shared__ int s_i[PC_THREADS / 32];
    __shared__ double s_lp[PC_MAXB];
    __shared__ double s_part[PC_THREADS / 32][PC_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // pivot = first index of the largest remaining diagonal entry (same result in every CTA)
    double v = -2.0;
    int idx = n;
    for (int i = tid; i < n; i += PC_THREADS) argmax_take(v, idx, d_in[i], i);
    for (int o = 16; o; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        argmax_take(v, idx, v2, i2);
    }
    if (lane == 0) { s_v[warp] = v; s_i[warp] = idx; }
    __syncthreads();
    v = s_v[0];
    idx = s_i[0];
#pragma unroll
    for (int w = 1; w < PC_THREADS / 32; ++w) argmax_take(v, idx, s_v[w], s_i[w]);

    const int row = blockIdx.x * PC_ROWS + lane;
    const long long colj = (long long)j * n;
    if (!(v > *thresh)) {
        // numerical rank reached (also on later columns: d is carried over unchanged)
        if (warp == 0 && row < n) {
            L[row + colj] = 0.0;
            d_out[row] = d_in[row];
        }
        if (blockIdx.x == 0 && tid == 0 && *rank == -1) *rank = j;
        return;
    }
    const int p = idx;
    const double piv = sqrt(v);
    const int nb = j - k0;                       // pending columns of this block
    for (int k = tid; k < nb; k += PC_THREADS) s_lp[k] = L[p + (long long)(k0 + k) * n];
    __syncthreads();
    double acc = 0.0;
    if (row < n)
        for (int k = warp; k < nb; k += PC_THREADS / 32)
            acc = fma(L[row + (long long)(k0 + k) * n], s_lp[k], acc);
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && row < n) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PC_THREADS / 32; ++w) t += s_part[w][lane];
        const double di = d_in[row];
        double l, dn;
        if (di < 0.0) {                          // an earlier pivot row: its Schur row is zero
            l = 0.0;
            dn = -1.0;
        } else if (row == p) {
            l = piv;
            dn = -1.0;
        } else {
            l = (S[row + (long long)p * n] - t) / piv;
            const double lim = sqrt(di);
            l = fmin(fmax(l, -lim), lim);
            dn = fmax(di - l * l, 0.0);
        }
        L[row + colj] = l;
        d_out[row] = dn;
    }
}

}  // namespace

// G: n x n symmetric PSD (ld n), not modified.  L: n x n (ld n); on return L L^T = G up to the
// stop threshold, columns >= rank are zero.  Returns the numerical rank.
long long psd_factor(Context* ctx, const double* G, long long n, double* L, int block) {
    TNR_CHECK(n >= 1 && n <= 46340, "psd_factor: n out of range");
    const int nb = std::max(1, std::min(block > 0 ? block : 64, PC_MAXB));
    const int ni = (int)n;
    double* S = dalloc(ctx, (size_t)n * n);
    double* d = dalloc(ctx, (size_t)2 * n + 1);
    double* thresh = d + 2 * n;
    int* d_rank = nullptr;
    TNR_CUDA(cudaMallocAsync((void**)&d_rank, sizeof(int), ctx->stream));
    TNR_CUDA(cudaMemcpyAsync(S, G, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice,
                             ctx->stream));
    pchol_init_kernel<<<1, 1024, 0, ctx->stream>>>(G, ni, d, thresh, d_rank);
    ctx->ctr.launches++;
    const unsigned grid = (unsigned)((n + PC_ROWS - 1) / PC_ROWS);
    int rank = -1;
    long long done = 0;
    for (long long k0 = 0; k0 < n && rank == -1; k0 += nb) {
        const long long jend = std::min<long long>(k0 + nb, n);
        for (long long j = k0; j < jend; ++j)
            pchol_column_kernel<<<grid, PC_THREADS, 0, ctx->stream>>>(
                S, L, ni, d + (j & 1) * n, d + ((j + 1) & 1) * n, (int)j, (int)k0, thresh, d_rank);
        ctx->ctr.launches += (unsigned long long)(jend - k0);
        TNR_CUDA(cudaGetLastError());
        done = jend;
        TNR_CUDA(cudaMemcpyAsync(&rank, d_rank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (rank == -1 && jend < n)
            gemm(ctx, 'N', 'T', ni, ni, (int)(jend - k0), -1.0, L + k0 * n, n, L + k0 * n, n, 1.0, S, n);
    }
    const bool finite = rank != -2;
    if (rank < 0) rank = ni;
    if (done < n)
        TNR_CUDA(cudaMemsetAsync(L + done * n, 0, (size_t)(n - done) * n * sizeof(double),
                                 ctx->stream));
    TNR_CUDA(cudaFreeAsync(d_rank, ctx->stream));
    dfree(ctx, d);
    dfree(ctx, S);
    ctx->ctr.psd_factorizations++;
    TNR_CHECK(finite, "psd_factor: non-finite diagonal entry");
    return rank;
}

// ---------------------------------------------------------------------------------------------
// CholeskyQR2: orthonormal basis of the columns of a tall, well-conditioned matrix with three
// kinds of kernels only -- Gram matrix (DMMA GEMM), a one-CTA Cholesky + triangular inverse of the
// small n x n matrix, and A R^-1 (DMMA GEMM) -- done twice, which brings ||Q^T Q - I|| from
// cond(A)^2 eps to eps.  It is what the block subspace iterations (the truncated SVDs of
// atrg3d.jl:35,43 on an implicit operator, and svd_topk / eigh_topk in tensor_ops.cu) run
// BETWEEN their Rayleigh-Ritz steps: those steps decide the result and keep the Householder QR +
// Jacobi path, the steps in between only have to keep the basis well conditioned.  Refuses
// (A untouched) when the Cholesky factor shows cond(A) >~ 1e5 or is not positive definite, or when
// A has more than 152 columns (the small matrix must fit in shared memory); the caller then
// takes the Householder path.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int CI_THREADS = 1024;
constexpr int CI_SMEM_N = 152;     // n <= 152: the matrix lives in shared memory (<= 185 KB)

// One CTA.  G (n x n, ld n, symmetric positive definite; upper triangle used) -> R in the upper
// triangle of G (G = R^T R), X = R^-1 (upper triangular, strict lower part zero).
// status[0] = 1 when a pivot is not positive / finite or min_k R_kk <= min_ratio * max_k R_kk.
__global__ void __launch_bounds__(CI_THREADS) chol_inv_kernel(double* __restrict__ G, int n,
                                                              double* __restrict__ X,
                                                              double min_ratio, int use_smem,
                                                              int* __restrict__ status) {
    extern __shared__ double ci_sm[];
    __shared__ double s_piv;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    double* M = use_smem ? ci_sm : G;
    if (use_smem)
        for (int e = tid; e < n * n; e += CI_THREADS) M[e] = G[e];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    double rmin = 1e300, rmax = 0.0;           // thread 0 only
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = M[k + (long long)k * n];
            if (!(d > 0.0) || !isfinite(d)) s_bad = 1;
            const double r = sqrt(d > 0.0 ? d : 1.0);
            s_piv = r;
            rmin = fmin(rmin, r);
            rmax = fmax(rmax, r);
        }
        __syncthreads();
        if (s_bad) break;
        const double piv = s_piv;
        for (int j = k + tid; j < n; j += CI_THREADS)
            M[k + (long long)j * n] = (j == k) ? piv : M[k + (long long)j * n] / piv;
        __syncthreads();
        const int w = n - k - 1;
        for (int e = tid; e < w * w; e += CI_THREADS) {
            const int i = k + 1 + e % w, j = k + 1 + e / w;
            if (i <= j) M[i + (long long)j * n] -= M[k + (long long)i * n] * M[k + (long long)j * n];
        }
        __syncthreads();
    }
    if (tid == 0 && (s_bad || !(rmin > min_ratio * rmax))) {
        s_bad = 1;
        status[0] = 1;
    }
    __syncthreads();
    if (s_bad) return;
    // X = R^-1 by back substitution, one thread per column
    for (int j = tid; j < n; j += CI_THREADS) {
        double* x = X + (long long)j * n;
        for (int i = n - 1; i > j; --i) x[i] = 0.0;
        x[j] = 1.0 / M[j + (long long)j * n];
        for (int i = j - 1; i >= 0; --i) {
            double sacc = 0.0;
            for (int k = i + 1; k <= j; ++k) sacc = fma(M[i + (long long)k * n], x[k], sacc);
            x[i] = -sacc / M[i + (long long)i * n];
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int e = tid; e < n * n; e += CI_THREADS) G[e] = M[e];
    }
}

void chol_inv(Context* ctx, double* G, long long n, double* X, double min_ratio, int* status) {
    const int use_smem = n <= CI_SMEM_N ? 1 : 0;
    const size_t smem = use_smem ? (size_t)n * n * sizeof(double) : 0;
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      CI_SMEM_N * CI_SMEM_N * (int)sizeof(double)));
        configured = true;
    }
    chol_inv_kernel<<<1, CI_THREADS, smem, ctx->stream>>>(G, (int)n, X, min_ratio, use_smem, status);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

}  // namespace

// A (m x n, ld m, m >= n) <- Q with orthonormal columns spanning the columns of A, Q = A R^-1 for
// the Cholesky factor R (positive diagonal) of A^T A.  Returns false and leaves A untouched when
// the factor is not safely positive definite (rank deficient or cond(A) >~ 1e5).
bool cholqr2(Context* ctx, double* A, long long m, long long n) {
    TNR_CHECK(m >= n && n >= 1, "cholqr2: expects a tall matrix");
    // beyond CI_SMEM_N columns the one-CTA factorization would run on global memory (measured:
    // TRG chi=128, 256 columns, 0.48 -> 0.86 s per step): those bases keep the Householder path
    if (n > CI_SMEM_N || m > 2147483647LL) return false;
    double* G = dalloc(ctx, (size_t)2 * n * n);
    double* X = G + n * n;
    double* tmp = dalloc(ctx, (size_t)m * n);
    int* d_status = nullptr;
    TNR_CUDA(cudaMallocAsync((void**)&d_status, 2 * sizeof(int), ctx->stream));
    TNR_CUDA(cudaMemsetAsync(d_status, 0, 2 * sizeof(int), ctx->stream));
    const int mi = (int)m, ni = (int)n;
    gemm(ctx, 'T', 'N', ni, ni, mi, 1.0, A, m, A, m, 0.0, G, n);
    chol_inv(ctx, G, n, X, 1e-5, d_status);
    gemm(ctx, 'N', 'N', mi, ni, ni, 1.0, A, m, X, n, 0.0, tmp, m);
    gemm(ctx, 'T', 'N', ni, ni, mi, 1.0, tmp, m, tmp, m, 0.0, G, n);
    chol_inv(ctx, G, n, X, 0.5, d_status + 1);        // second pass: R ~ I
    int st[2] = {1, 1};
    TNR_CUDA(cudaMemcpyAsync(st, d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TNR_CUDA(cudaStreamSynchronize(ctx->stream));
    const bool ok = st[0] == 0 && st[1] == 0;
    if (ok) gemm(ctx, 'N', 'N', mi, ni, ni, 1.0, tmp, m, X, n, 0.0, A, m);
    TNR_CUDA(cudaFreeAsync(d_status, ctx->stream));
    dfree(ctx, tmp);
    dfree(ctx, G);
    if (ok) ctx->ctr.cholqr2++; else ctx->ctr.cholqr2_refused++;
    return ok;
}

}  // namespace tnr
