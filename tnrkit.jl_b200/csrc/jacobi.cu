// One-sided (Hestenes) Jacobi orthogonalisation: the engine behind svd_trunc,
// eigh_trunc and the R-factor of left_orth/right_orth on the device.
//
// Replaces the LAPACK calls behind MatrixAlgebraKit `svd_trunc` / `eigh_trunc!` /
// `left_orth` used by the reference (src/utility/projectors.jl:213-219,
// src/schemes/btrg.jl:63, hotrg.jl:106,114, atrg3d.jl:37,58-66).
//
// Column pairs follow a round-robin tournament: every round rotates n/2 disjoint
// pairs in parallel (one CTA per pair, warp-shuffle reductions for the three dot
// products), n-1 rounds per sweep.  Convergence is checked per sweep through a
// device counter of applied rotations.  Selection of the chi largest values and
// the truncation error are computed on the device (rank-by-counting), so the only
// host round trip per factorisation is the per-sweep convergence flag.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tensor.hpp"

namespace tnr {
namespace {

// A rotation is SIGNIFICANT when the off-diagonal Gram entry it removes exceeds 1e-13 ||A||_F^2,
// i.e. it matters at the absolute accuracy (relative to sigma_1) that a LAPACK SVD / eigh gives.
// A sweep limit reached with only insignificant rotations left -- the relative criterion still
// firing among columns of norm << sigma_1, typical of graded spectra -- is accepted and counted;
// a limit reached with significant rotations left is an error.
__device__ constexpr double BIG_ROT = 1e-13 * 1e34;   // times floor2 = 1e-34 ||A||_F^2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NT>
__global__ void __launch_bounds__(NT) jacobi_round_kernel(double* __restrict__ G, long long m,
                                                          long long n, long long ldg,
                                                          double* __restrict__ V, long long ldv,
                                                          int round, int npad, double tol,
                                                          double floor2, int* rot_count) {
    int i = blockIdx.x;
    int a, b;
    if (i == 0) {
        a = npad - 1;
        b = round;
    } else {
        a = (round + i) % (npad - 1);
        b = (round - i + npad - 1) % (npad - 1);
    }
    int p = min(a, b), q = max(a, b);
    if (q >= n) return;
    double* gp = G + (long long)p * ldg;
    double* gq = G + (long long)q * ldg;
    double al = 0.0, be = 0.0, ga = 0.0;
    for (long long r = threadIdx.x; r < m; r += NT) {
        double x = gp[r], y = gq[r];
        al += x * x;
        be += y * y;
        ga += x * y;
    }
    __shared__ double red[3][NT / 32];
    __shared__ double cs[2];
    al = warp_sum(al);
    be = warp_sum(be);
    ga = warp_sum(ga);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = al;
        red[1][warp] = be;
        red[2][warp] = ga;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A = 0, B = 0, C = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            A += red[0][w];
            B += red[1][w];
            C += red[2][w];
        }
        double c = 1.0, s = 0.0;
        bool rotate = (fabs(C) > tol * sqrt(A * B)) && (A > floor2) && (B > floor2);
        if (rotate) {
            double zeta = (B - A) / (2.0 * C);
            double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            c = 1.0 / sqrt(1.0 + t * t);
            s = c * t;
            atomicAdd(rot_count, 1);
            if (fabs(C) > BIG_ROT * floor2) atomicAdd(rot_count + 1, 1);
        }
        cs[0] = c;
        cs[1] = s;
    }
    __syncthreads();
    double c = cs[0], s = cs[1];
    if (s == 0.0) return;
    for (long long r = threadIdx.x; r < m; r += NT) {
        double x = gp[r], y = gq[r];
        gp[r] = c * x - s * y;
        gq[r] = s * x + c * y;
    }
    if (V) {
        double* vp = V + (long long)p * ldv;
        double* vq = V + (long long)q * ldv;
        for (long long r = threadIdx.x; r < n; r += NT) {
            double x = vp[r], y = vq[r];
            vp[r] = c * x - s * y;
            vq[r] = s * x + c * y;
        }
    }
}

// Block variant for matrices whose column pairs fit in shared memory: one CTA owns a PAIR OF
// COLUMN BLOCKS (2*BC columns of G and of V), performs a complete cyclic sweep over all
// 2BC(2BC-1)/2 column pairs inside shared memory (inner round-robin, one warp per pair, warp
// shuffle reductions for the three dot products) and writes the columns back.  A sweep over the
// matrix then needs n/BC - 1 launches instead of n - 1, and every rotation works on shared
// memory instead of L2.
template <int BC>
__device__ __forceinline__ void jacobi_block_round_body(
    double* __restrict__ G, int m, int n, long long ldg, double* __restrict__ V, int nv,
    long long ldv, int round, int nblk_pad, double tol, double floor2, int* rot_count) {
    extern __shared__ double sm[];
    constexpr int NC = 2 * BC;
    const int rows = m + (V ? nv : 0);       // G column followed by V column
    double* cols = sm;                        // [NC][rows]
    __shared__ int colid[NC];
    int i = blockIdx.x, a, b;
    if (i == 0) { a = nblk_pad - 1; b = round; }
    else { a = (round + i) % (nblk_pad - 1); b = (round - i + nblk_pad - 1) % (nblk_pad - 1); }
    const int bp = min(a, b), bq = max(a, b);
    if (threadIdx.x < NC) {
        int c = threadIdx.x < BC ? bp * BC + threadIdx.x : bq * BC + (threadIdx.x - BC);
        colid[threadIdx.x] = (c < n) ? c : -1;
    }
    __syncthreads();
    // load
    for (int c = 0; c < NC; ++c) {
        int gc = colid[c];
        if (gc < 0) continue;
        const double* g = G + (long long)gc * ldg;
        double* d = cols + (long long)c * rows;
        // L2 loads: in the persistent kernel another SM wrote these columns in the last round
        for (int r = threadIdx.x; r < m; r += 256) d[r] = __ldcg(g + r);
        if (V) {
            const double* v = V + (long long)gc * ldv;
            for (int r = threadIdx.x; r < nv; r += 256) d[m + r] = __ldcg(v + r);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int rotated = 0, big = 0;
    for (int rr = 0; rr < NC - 1; ++rr) {
        for (int pi = warp; pi < BC; pi += 8) {
            int x, y;
            if (pi == 0) { x = NC - 1; y = rr; }
            else { x = (rr + pi) % (NC - 1); y = (rr - pi + NC - 1) % (NC - 1); }
            int p = min(x, y), q = max(x, y);
            if (colid[p] < 0 || colid[q] < 0) continue;
            double* cp = cols + (long long)p * rows;
            double* cq = cols + (long long)q * rows;
            double al = 0.0, be = 0.0, ga = 0.0;
            for (int r = lane; r < m; r += 32) {
                double u = cp[r], w = cq[r];
                al += u * u; be += w * w; ga += u * w;
            }
            al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
            bool rot = (fabs(ga) > tol * sqrt(al * be)) && (al > floor2) && (be > floor2);
            if (rot) {
                double zeta = (be - al) / (2.0 * ga);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = lane; r < rows; r += 32) {
                    double u = cp[r], w = cq[r];
                    cp[r] = c * u - s * w;
                    cq[r] = s * u + c * w;
                }
                if (lane == 0) {
                    ++rotated;
                    if (fabs(ga) > BIG_ROT * floor2) ++big;
                }
            }
        }
        __syncthreads();
    }
    if (lane == 0 && rotated) atomicAdd(rot_count, rotated);
    if (lane == 0 && big) atomicAdd(rot_count + 1, big);
    // store
    for (int c = 0; c < NC; ++c) {
        int gc = colid[c];
        if (gc < 0) continue;
        double* g = G + (long long)gc * ldg;
        const double* d = cols + (long long)c * rows;
        for (int r = threadIdx.x; r < m; r += 256) g[r] = d[r];
        if (V) {
            double* v = V + (long long)gc * ldv;
            for (int r = threadIdx.x; r < nv; r += 256) v[r] = d[m + r];
        }
    }
}

template <int BC>
__global__ void __launch_bounds__(256) jacobi_block_round_kernel(
    double* __restrict__ G, int m, int n, long long ldg, double* __restrict__ V, int nv,
    long long ldv, int round, int nblk_pad, double tol, double floor2, int* rot_count) {
    jacobi_block_round_body<BC>(G, m, n, ldg, V, nv, ldv, round, nblk_pad, tol, floor2, rot_count);
}

// The whole iteration in ONE cooperative launch: all nblk_pad/2 CTAs stay resident, a round ends
// with a grid barrier (monotonic atomic counter), a sweep ends with every CTA reading the
// rotation counter of that sweep; the host reads (sweeps, converged) once at the end.  Replaces
// ~(nblk-1) launches + one host synchronisation PER SWEEP of the small dense eigenproblems /
// SVDs (Ritz problems of the subspace solvers, R factors of the QR stage), which were more than
// half of the device time of the SVD-bound steps (profiles/r02_share_*.md).
__device__ __forceinline__ void jacobi_grid_barrier(unsigned* bar, unsigned nblocks,
                                                    unsigned& phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned target = (++phase) * nblocks;
        atomicAdd(bar, 1u);
        while (*((volatile unsigned*)bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <int BC>
__global__ void __launch_bounds__(256) jacobi_block_persistent_kernel(
    double* __restrict__ G, int m, int n, long long ldg, double* __restrict__ V, int nv,
    long long ldv, int nblk_pad, double tol, double floor2, int max_sweeps,
    int* rot_counts /*[2 * (max_sweeps + 1)]*/, unsigned* bar,
    int* result /*[3]: sweeps, converged, significant rotations of the last sweep*/) {
    unsigned phase = 0;
    int sweeps = 0, converged = 0, last_big = 0;
    for (; sweeps < max_sweeps; ++sweeps) {
        for (int round = 0; round < nblk_pad - 1; ++round) {
            jacobi_block_round_body<BC>(G, m, n, ldg, V, nv, ldv, round, nblk_pad, tol, floor2,
                                        rot_counts + 2 * sweeps);
            jacobi_grid_barrier(bar, gridDim.x, phase);
        }
        const int rot = *((volatile int*)(rot_counts + 2 * sweeps));
        last_big = *((volatile int*)(rot_counts + 2 * sweeps + 1));
        if (rot == 0) { ++sweeps; converged = 1; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        result[0] = sweeps; result[1] = converged; result[2] = last_big;
    }
}

// Convergence test of the one-sided iteration from the Gram matrix: counts the column pairs
// that still violate |g_i . g_j| <= tol ||g_i|| ||g_j|| (same criterion as the sweeps).
__global__ void gram_violations_kernel(const double* __restrict__ gram, int n, double tol,
                                       double floor2, int* count) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    int i = (int)(idx % n), j = (int)(idx / n);
    if (i >= j) return;
    double a = gram[(long long)i * n + i], b = gram[(long long)j * n + j], c = gram[(long long)j * n + i];
    if (fabs(c) > tol * sqrt(a * b) && a > floor2 && b > floor2) atomicAdd(count, 1);
}

__global__ void __launch_bounds__(256) frob2_kernel(const double* __restrict__ G, long long m,
                                                    long long n, long long ldg, double* out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long long idx = threadIdx.x; idx < m * n; idx += 256) {
        double x = G[(idx / m) * ldg + idx % m];
        acc += x * x;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

__global__ void __launch_bounds__(128) column_values_kernel(const double* __restrict__ G,
                                                            long long m, long long ldg,
                                                            const double* __restrict__ V,
                                                            long long ldv, double* vals,
                                                            int signed_rayleigh) {
    int j = blockIdx.x;
    const double* g = G + (long long)j * ldg;
    double acc = 0.0;
    if (signed_rayleigh) {
        const double* v = V + (long long)j * ldv;
        for (long long r = threadIdx.x; r < m; r += 128) acc += g[r] * v[r];
    } else {
        for (long long r = threadIdx.x; r < m; r += 128) acc += g[r] * g[r];
    }
    __shared__ double red[4];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = red[0] + red[1] + red[2] + red[3];
        vals[j] = signed_rayleigh ? s : sqrt(s);
    }
}

__global__ void rank_kernel(const double* __restrict__ vals, long long n, int* rank) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    double vj = fabs(vals[j]);
    int r = 0;
    for (long long i = 0; i < n; ++i) {
        double vi = fabs(vals[i]);
        r += (vi > vj) || (vi == vj && i < j);
    }
    rank[j] = r;
}

__global__ void __launch_bounds__(1024) trunc_eps_kernel(const double* __restrict__ vals,
                                                         const int* __restrict__ rank,
                                                         long long n, long long k, double* eps) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (long long j = threadIdx.x; j < n; j += 1024)
        if (rank[j] >= k) acc += vals[j] * vals[j];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *eps = sqrt(red[0]);
}

__global__ void gather_columns_kernel(const double* __restrict__ src, long long m, long long lds,
                                      const int* __restrict__ rank, long long k, double* dst,
                                      long long ldd, const double* vals, int normalize) {
    int j = blockIdx.y;
    int r = rank[j];
    if (r >= k) return;
    double f = 1.0;
    if (normalize) {
        double s = fabs(vals[j]);
        f = (s > 0.0) ? 1.0 / s : 0.0;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
         i += (long long)gridDim.x * blockDim.x)
        dst[(long long)r * ldd + i] = f * src[(long long)j * lds + i];
}

__global__ void gather_values_kernel(const double* vals, long long n, const int* rank, long long k,
                                     double* out, int absval) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    int r = rank[j];
    if (r < k) out[r] = absval ? fabs(vals[j]) : vals[j];
}

}  // namespace

int jacobi_orthogonalize(Context* ctx, double* G, long long m, long long n, long long ldg,
                         double* V, long long ldv) {
    if (n <= 1 || m <= 0) return 0;
    const int npad = (int)((n + 1) & ~1LL);
    const double eps = 2.220446049250313e-16;
    const double tol = std::max(1e-15, std::sqrt((double)m) * eps);
    // columns whose squared norm is below (1e-17 ||A||_F)^2 are numerical zeros
    double* d_f2 = dalloc(ctx, 1);
    int* d_rot;
    TNR_CUDA(cudaMallocAsync((void**)&d_rot, 2 * sizeof(int), ctx->stream));
    if (ldg == m) sum_squares(ctx, G, m * n, d_f2);
    else frob2_kernel<<<1, 256, 0, ctx->stream>>>(G, m, n, ldg, d_f2);
    ctx->ctr.launches++;
    double f2 = 0.0;
    TNR_CUDA(cudaMemcpyAsync(&f2, d_f2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TNR_CUDA(cudaStreamSynchronize(ctx->stream));
    const double floor2 = f2 * 1e-34;
    const int max_sweeps = ctx->jacobi_max_sweeps;
    int sweeps = 0;
    // Tall matrices: precondition with the eigenvectors W of the Gram matrix G^T G (one DMMA
    // GEMM + a small n x n Jacobi) and rotate G <- G W, V <- V W by GEMM.  G W already has
    // nearly orthogonal columns, so the accurate one-sided sweeps below converge in 1-2 passes
    // over the tall matrix instead of ~10; accuracy is that of Jacobi on G W (W is orthogonal).
    // Tall matrices, the design the north star names: blocked Householder QR (trailing update on
    // the tensor cores, qr.cu), one-sided Jacobi on the small n x n factor R in shared memory
    // (the block kernel below, via the recursive call), then G <- Q (R V).  The rotations only
    // ever touch R; the tall matrix is read by the QR panels and written once by the GEMMs
    // that apply Q.
    if (!ctx->disable_qr && m >= 2 * n && n >= 32 && ldg == m && (!V || ldv == n)) {
        TNR_CUDA(cudaFreeAsync(d_rot, ctx->stream));
        dfree(ctx, d_f2);
        QRWork w;
        qr_factor(ctx, G, m, n, ldg, w);
        double* R = dalloc(ctx, (size_t)n * n);
        qr_copy_r(ctx, G, ldg, n, R, n);
        // R is preconditioned by the eigenvectors W of R^T R (= A^T A, but an n^3 product) before
        // the sweeps: one-sided Jacobi on the columns of a bare upper-triangular, graded R
        // converges slowly, on R W it needs one or two sweeps (same scheme as the round-1 tall
        // path, now entirely on n x n matrices)
        if (!ctx->disable_precondition) {
            double* gram = dalloc(ctx, (size_t)n * n);
            double* W = dalloc(ctx, (size_t)n * n);
            double* tmp = dalloc(ctx, (size_t)n * n);
            gemm(ctx, 'T', 'N', (int)n, (int)n, (int)n, 1.0, R, n, R, n, 0.0, gram, n);
            symmetrize(ctx, gram, n);
            set_identity(ctx, W, n);
            jacobi_orthogonalize(ctx, gram, n, n, n, W, n);
            gemm(ctx, 'N', 'N', (int)n, (int)n, (int)n, 1.0, R, n, W, n, 0.0, tmp, n);
            TNR_CUDA(cudaMemcpyAsync(R, tmp, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
            if (V) {
                gemm(ctx, 'N', 'N', (int)n, (int)n, (int)n, 1.0, V, ldv, W, n, 0.0, tmp, n);
                TNR_CUDA(cudaMemcpy2DAsync(V, ldv * sizeof(double), tmp, n * sizeof(double),
                                           n * sizeof(double), n, cudaMemcpyDeviceToDevice,
                                           ctx->stream));
            }
            dfree(ctx, gram);
            dfree(ctx, W);
            dfree(ctx, tmp);
        }
        const int sw = jacobi_orthogonalize(ctx, R, n, n, n, V, ldv);   // R <- R V_J
        if (std::getenv("TNR_TRACE"))
            fprintf(stderr, "[jacobi via QR] %lld x %lld: %d sweeps on R\n", m, n, sw);
        qr_q_times(ctx, w, R, n, n, G, ldg);                             // G <- Q [R V_J; 0]
        dfree(ctx, R);
        return sw;
    }
    if (!ctx->disable_precondition && m >= 4 * n && n >= 32 && ldg == m && (!V || ldv == n)) {
        double* gram = dalloc(ctx, (size_t)n * n);
        double* W = dalloc(ctx, (size_t)n * n);
        gemm(ctx, 'T', 'N', (int)n, (int)n, (int)m, 1.0, G, ldg, G, ldg, 0.0, gram, n);
        symmetrize(ctx, gram, n);
        set_identity(ctx, W, n);
        jacobi_orthogonalize(ctx, gram, n, n, n, W, n);
        double* tmp = dalloc(ctx, (size_t)m * n);
        gemm(ctx, 'N', 'N', (int)m, (int)n, (int)n, 1.0, G, ldg, W, n, 0.0, tmp, m);
        TNR_CUDA(cudaMemcpyAsync(G, tmp, (size_t)m * n * sizeof(double), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
        dfree(ctx, tmp);
        if (V) {
            double* tv = dalloc(ctx, (size_t)n * n);
            gemm(ctx, 'N', 'N', (int)n, (int)n, (int)n, 1.0, V, ldv, W, n, 0.0, tv, n);
            TNR_CUDA(cudaMemcpyAsync(V, tv, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
            dfree(ctx, tv);
        }
        dfree(ctx, gram);
        dfree(ctx, W);
        ctx->ctr.preconditioned_jacobi++;
    }
    // block variant when a pair of column blocks (G and V columns) fits in shared memory
    const long long rows = m + (V ? n : 0);
    int BC = 0;
    // Time of a sweep ~ (n / BC rounds) x (2 BC - 1 inner rounds) x ceil(BC / 8 warps) pair
    // rotations: BC = 16 runs two passes of the 8 warps per inner round on half as many CTAs as
    // BC = 8 (n = 256: 15 x 62 against 31 x 15 rotation times per sweep, measured 13 ms per
    // Ritz problem on 8 CTAs in profiles/r02_share_trg128_potts_qr.md), so 8 is the default cap.
    for (int cand : {16, 8, 4})
        if (!BC && cand <= ctx->jacobi_max_bc && 2LL * cand * rows * 8 <= 200 * 1024 &&
            n >= 2 * cand) BC = cand;
    if (BC && !ctx->disable_block_jacobi) {
        const int nblk = (int)((n + BC - 1) / BC);
        const int nblk_pad = (nblk + 1) & ~1;
        const size_t smem = (size_t)2 * BC * rows * 8;
        auto launch = [&](int round) {
            if (BC == 16) jacobi_block_round_kernel<16><<<nblk_pad / 2, 256, smem, ctx->stream>>>(
                G, (int)m, (int)n, ldg, V, (int)n, ldv, round, nblk_pad, tol, floor2, d_rot);
            else if (BC == 8) jacobi_block_round_kernel<8><<<nblk_pad / 2, 256, smem, ctx->stream>>>(
                G, (int)m, (int)n, ldg, V, (int)n, ldv, round, nblk_pad, tol, floor2, d_rot);
            else jacobi_block_round_kernel<4><<<nblk_pad / 2, 256, smem, ctx->stream>>>(
                G, (int)m, (int)n, ldg, V, (int)n, ldv, round, nblk_pad, tol, floor2, d_rot);
        };
        static bool configured = false;
        if (!configured) {
            TNR_CUDA(cudaFuncSetAttribute(jacobi_block_round_kernel<16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            TNR_CUDA(cudaFuncSetAttribute(jacobi_block_round_kernel<8>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            TNR_CUDA(cudaFuncSetAttribute(jacobi_block_round_kernel<4>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured = true;
        }
        bool converged = false;
        int last_big = 0;
        if (!ctx->disable_persistent_jacobi && nblk_pad / 2 <= ctx->num_sms) {
            // one cooperative launch for the whole iteration
            static bool pconf = false;
            if (!pconf) {
                TNR_CUDA(cudaFuncSetAttribute(jacobi_block_persistent_kernel<16>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                TNR_CUDA(cudaFuncSetAttribute(jacobi_block_persistent_kernel<8>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                TNR_CUDA(cudaFuncSetAttribute(jacobi_block_persistent_kernel<4>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                pconf = true;
            }
            int* d_state = nullptr;    // [max_sweeps + 1] rotation counters, barrier word, result[2]
            const size_t nstate = 2 * ((size_t)max_sweeps + 1) + 1 + 3;
            TNR_CUDA(cudaMallocAsync((void**)&d_state, nstate * sizeof(int), ctx->stream));
            TNR_CUDA(cudaMemsetAsync(d_state, 0, nstate * sizeof(int), ctx->stream));
            int* d_rots = d_state;
            unsigned* d_bar = reinterpret_cast<unsigned*>(d_state + 2 * (max_sweeps + 1));
            int* d_res = d_state + 2 * (max_sweeps + 1) + 1;
            int mi = (int)m, ni = (int)n, nvi = (int)n, npad_blk = nblk_pad, ms = max_sweeps;
            double tol_ = tol, floor_ = floor2;
            void* args[] = {&G, &mi, &ni, &ldg, &V, &nvi, &ldv, &npad_blk, &tol_, &floor_, &ms,
                            &d_rots, &d_bar, &d_res};
            const void* fn = BC == 16 ? (const void*)jacobi_block_persistent_kernel<16>
                           : BC == 8  ? (const void*)jacobi_block_persistent_kernel<8>
                                      : (const void*)jacobi_block_persistent_kernel<4>;
            TNR_CUDA(cudaLaunchCooperativeKernel(fn, dim3(nblk_pad / 2), dim3(256), args, smem,
                                                 ctx->stream));
            ctx->ctr.launches++;
            ctx->ctr.persistent_jacobi++;
            int res[3] = {0, 0, 0};
            TNR_CUDA(cudaMemcpyAsync(res, d_res, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            TNR_CUDA(cudaStreamSynchronize(ctx->stream));
            TNR_CUDA(cudaFreeAsync(d_state, ctx->stream));
            sweeps = res[0];
            converged = res[1] != 0;
            last_big = res[2];
        } else
        for (; sweeps < max_sweeps; ++sweeps) {
            TNR_CUDA(cudaMemsetAsync(d_rot, 0, 2 * sizeof(int), ctx->stream));
            for (int round = 0; round < nblk_pad - 1; ++round) launch(round);
            ctx->ctr.launches += nblk_pad - 1;
            TNR_CUDA(cudaGetLastError());
            int rot[2] = {0, 0};
            TNR_CUDA(cudaMemcpyAsync(rot, d_rot, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            TNR_CUDA(cudaStreamSynchronize(ctx->stream));
            last_big = rot[1];
            if (rot[0] == 0) { ++sweeps; converged = true; break; }
        }
        TNR_CUDA(cudaFreeAsync(d_rot, ctx->stream));
        dfree(ctx, d_f2);
        // a factorisation that did not converge must not flow silently into U / S / V; a limit
        // reached with only rounding-level rotations left (|g_i.g_j| <= 1e-9 |g_i||g_j|) is
        // accepted and counted
        if (!converged && last_big == 0) { converged = true; ctx->ctr.jacobi_limit_accepted++; }
        if (!converged) ctx->ctr.jacobi_not_converged++;
        TNR_CHECK(converged, "one-sided Jacobi did not converge within the sweep limit (" +
                                 std::to_string(m) + " x " + std::to_string(n) + ")");
        return sweeps;
    }
    // tall problems: a Gram GEMM (2 m n^2 flop on the tensor cores) is far cheaper than a sweep
    // (n-1 passes over the m x n matrix), so convergence is tested on the Gram matrix before
    // every sweep -- a preconditioned problem usually needs zero or one sweep.
    const bool gram_check = !ctx->disable_precondition && m >= 4 * n && n >= 32 && ldg == m;
    bool converged = false;
    int last_big = 0;
    for (; sweeps < max_sweeps; ++sweeps) {
        if (gram_check) {
            double* gram = dalloc(ctx, (size_t)n * n);
            gemm(ctx, 'T', 'N', (int)n, (int)n, (int)m, 1.0, G, ldg, G, ldg, 0.0, gram, n);
            TNR_CUDA(cudaMemsetAsync(d_rot, 0, 2 * sizeof(int), ctx->stream));
            long long nn = n * n;
            gram_violations_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, ctx->stream>>>(
                gram, (int)n, tol, floor2, d_rot);
            ctx->ctr.launches++;
            int viol = 0;
            TNR_CUDA(cudaMemcpyAsync(&viol, d_rot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            TNR_CUDA(cudaStreamSynchronize(ctx->stream));
            dfree(ctx, gram);
            if (viol == 0) { converged = true; break; }
        }
        TNR_CUDA(cudaMemsetAsync(d_rot, 0, 2 * sizeof(int), ctx->stream));
        for (int round = 0; round < npad - 1; ++round) {
            if (m > 2048)
                jacobi_round_kernel<256><<<npad / 2, 256, 0, ctx->stream>>>(
                    G, m, n, ldg, V, ldv, round, npad, tol, floor2, d_rot);
            else
                jacobi_round_kernel<128><<<npad / 2, 128, 0, ctx->stream>>>(
                    G, m, n, ldg, V, ldv, round, npad, tol, floor2, d_rot);
        }
        ctx->ctr.launches += npad - 1;
        TNR_CUDA(cudaGetLastError());
        int rot[2] = {0, 0};
        TNR_CUDA(cudaMemcpyAsync(rot, d_rot, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->stream));
        last_big = rot[1];
        if (rot[0] == 0) {
            ++sweeps;
            converged = true;
            break;
        }
    }
    TNR_CUDA(cudaFreeAsync(d_rot, ctx->stream));
    if (!converged && last_big == 0 && sweeps > 0) { converged = true; ctx->ctr.jacobi_limit_accepted++; }
    if (!converged) {
        ctx->ctr.jacobi_not_converged++;
        dfree(ctx, d_f2);
        TNR_CHECK(false, "one-sided Jacobi did not converge within the sweep limit (" + std::to_string(m) +
                             " x " + std::to_string(n) + ")");
    }
    dfree(ctx, d_f2);
    return sweeps;
}

void column_values(Context* ctx, const double* G, long long m, long long n, long long ldg,
                   const double* V, long long ldv, double* vals, bool signed_rayleigh) {
    if (n <= 0) return;
    column_values_kernel<<<(unsigned)n, 128, 0, ctx->stream>>>(G, m, ldg, V, ldv, vals,
                                                               signed_rayleigh ? 1 : 0);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void rank_select(Context* ctx, const double* vals, long long n, long long k, int* rank,
                 double* dev_eps) {
    rank_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(vals, n, rank);
    if (dev_eps) trunc_eps_kernel<<<1, 1024, 0, ctx->stream>>>(vals, rank, n, k, dev_eps);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches += dev_eps ? 2 : 1;
}

void gather_columns(Context* ctx, const double* src, long long m, long long n, long long lds,
                    const int* rank, long long k, double* dst, long long ldd, const double* vals,
                    bool normalize) {
    if (n <= 0 || m <= 0) return;
    unsigned gx = (unsigned)std::min<long long>((m + 255) / 256, 64);
    TNR_CHECK(n <= 65535, "gather_columns: too many columns");
    dim3 grid(gx, (unsigned)n);
    gather_columns_kernel<<<grid, 256, 0, ctx->stream>>>(src, m, lds, rank, k, dst, ldd, vals,
                                                         normalize ? 1 : 0);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void gather_values(Context* ctx, const double* vals, long long n, const int* rank, long long k,
                   double* out, bool absval) {
    gather_values_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(vals, n, rank, k, out,
                                                                             absval ? 1 : 0);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

}  // namespace tnr
