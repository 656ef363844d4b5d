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

}  // namespace tnr
