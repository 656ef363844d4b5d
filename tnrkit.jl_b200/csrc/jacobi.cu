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

#include "common.cuh"

namespace tnr {
namespace {

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
    TNR_CUDA(cudaMallocAsync((void**)&d_rot, sizeof(int), ctx->stream));
    frob2_kernel<<<1, 256, 0, ctx->stream>>>(G, m, n, ldg, d_f2);
    ctx->ctr.launches++;
    double f2 = 0.0;
    TNR_CUDA(cudaMemcpyAsync(&f2, d_f2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TNR_CUDA(cudaStreamSynchronize(ctx->stream));
    const double floor2 = f2 * 1e-34;
    const int max_sweeps = 40;
    int sweeps = 0;
    for (; sweeps < max_sweeps; ++sweeps) {
        TNR_CUDA(cudaMemsetAsync(d_rot, 0, sizeof(int), ctx->stream));
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
        int rot = 0;
        TNR_CUDA(cudaMemcpyAsync(&rot, d_rot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (rot == 0) {
            ++sweeps;
            break;
        }
    }
    TNR_CUDA(cudaFreeAsync(d_rot, ctx->stream));
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
