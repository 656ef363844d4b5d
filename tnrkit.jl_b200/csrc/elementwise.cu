// Small element-wise / reduction kernels: normalisation (finalize!), bond-weight
// scaling (U*sqrt(S), pseudopow), hermitian projection, projector selection.
// Reference: /root/reference/src/utility/finalize.jl:4-14,56-66,
// src/schemes/btrg.jl:51-60, src/schemes/hotrg.jl:106-118.
#include "common.cuh"
#include <algorithm>

namespace tnr {
namespace {

constexpr double PSEUDOPOW_TOL = 1.8189894035458565e-12;  // eps(Float64)^(3/4)

__device__ __forceinline__ double apply_mode(double s, int mode, double p) {
    if (mode == 1) return sqrt(s);
    if (mode == 2) return (s < PSEUDOPOW_TOL) ? s : pow(s, p);
    return s;
}

__global__ void scale_kernel(double* x, long long n, double alpha, const double* dev_div) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double a = dev_div ? 1.0 / *dev_div : alpha;
    if (i < n) x[i] *= a;
}

__global__ void diag_scale_kernel(double* A, long long m, long long n, long long lda,
                                  const double* s, int rows, int mode, double p) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m * n) return;
    long long r = i % m, c = i / m;
    double f = apply_mode(s[rows ? r : c], mode, p);
    A[c * lda + r] *= f;
}

__global__ void axis_scale_kernel(double* A, long long m1, long long n, long long total,
                                  const double* s, int mode, double p) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long j = (i / m1) % n;
    A[i] *= apply_mode(s[j], mode, p);
}

__global__ void vec_map_kernel(const double* s, double* out, long long n, int mode, double p) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = apply_mode(s[i], mode, p);
}

struct SumParams {
    int rank;
    long long dims[4];
    long long stride[4];
    const double* w[4];
    long long total;
};

// single block; fixed reduction tree -> deterministic
__global__ void __launch_bounds__(1024) strided_sum_kernel(const double* __restrict__ src,
                                                           SumParams p, double* out, int absval) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (long long idx = threadIdx.x; idx < p.total; idx += blockDim.x) {
        long long rest = idx, off = 0;
        double w = 1.0;
        for (int d = 0; d < p.rank; ++d) {
            long long i = rest % p.dims[d];
            rest /= p.dims[d];
            off += i * p.stride[d];
            if (p.w[d]) w *= p.w[d][i];
        }
        acc += w * src[off];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = absval ? fabs(red[0]) : red[0];
}

__global__ void symmetrize_kernel(double* A, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    long long r = i % n, c = i / n;
    if (r < c) {
        double v = 0.5 * (A[c * n + r] + A[r * n + c]);
        A[c * n + r] = v;
        A[r * n + c] = v;
    }
}

__global__ void identity_kernel(double* A, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    A[i] = (i % n == i / n) ? 1.0 : 0.0;
}

__global__ void select_copy_kernel(double* dst, const double* a, const double* b, long long n,
                                   const double* ea, const double* eb, double* eout) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    bool use_b = (*ea > *eb);
    if (i < n) dst[i] = use_b ? b[i] : a[i];
    if (i == 0 && eout) *eout = use_b ? *eb : *ea;
}

inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void scale(Context* ctx, double* x, long long n, double alpha) {
    if (n <= 0) return;
    scale_kernel<<<nblk(n), 256, 0, ctx->stream>>>(x, n, alpha, nullptr);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void scale_inv_dev(Context* ctx, double* x, long long n, const double* dev_scalar) {
    if (n <= 0) return;
    scale_kernel<<<nblk(n), 256, 0, ctx->stream>>>(x, n, 1.0, dev_scalar);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void diag_scale(Context* ctx, double* A, long long m, long long n, long long lda, const double* s,
                bool rows, int mode, double p) {
    if (m * n <= 0) return;
    diag_scale_kernel<<<nblk(m * n), 256, 0, ctx->stream>>>(A, m, n, lda, s, rows ? 1 : 0, mode, p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void axis_scale(Context* ctx, double* A, long long m1, long long n, long long m2, const double* s,
                int mode, double p) {
    long long total = m1 * n * m2;
    if (total <= 0) return;
    axis_scale_kernel<<<nblk(total), 256, 0, ctx->stream>>>(A, m1, n, total, s, mode, p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void vec_map(Context* ctx, const double* s, double* out, long long n, int mode, double p) {
    if (n <= 0) return;
    vec_map_kernel<<<nblk(n), 256, 0, ctx->stream>>>(s, out, n, mode, p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void strided_sum_w(Context* ctx, const double* src, int rank, const long long* dims,
                   const long long* stride, const double* const* weights, double* dev_out,
                   bool absval) {
    TNR_CHECK(rank >= 1 && rank <= 4, "strided_sum: rank must be 1..4");
    SumParams p{};
    p.rank = rank;
    p.total = 1;
    for (int d = 0; d < rank; ++d) {
        p.dims[d] = dims[d];
        p.stride[d] = stride[d];
        p.w[d] = weights ? weights[d] : nullptr;
        p.total *= dims[d];
    }
    strided_sum_kernel<<<1, 1024, 0, ctx->stream>>>(src, p, dev_out, absval ? 1 : 0);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void strided_sum(Context* ctx, const double* src, int rank, const long long* dims,
                 const long long* stride, double* dev_out, bool absval) {
    strided_sum_w(ctx, src, rank, dims, stride, nullptr, dev_out, absval);
}

void symmetrize(Context* ctx, double* A, long long n) {
    symmetrize_kernel<<<nblk(n * n), 256, 0, ctx->stream>>>(A, n);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void set_identity(Context* ctx, double* A, long long n) {
    identity_kernel<<<nblk(n * n), 256, 0, ctx->stream>>>(A, n);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void fill_zero(Context* ctx, double* A, long long n) {
    TNR_CUDA(cudaMemsetAsync(A, 0, n * sizeof(double), ctx->stream));
}

__global__ void zero_small_columns_kernel(double* A, long long m, long long lda,
                                          const double* __restrict__ vals, double rel) {
    const int j = blockIdx.y;
    if (!(fabs(vals[j]) <= rel * fabs(vals[0]))) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m;
         i += (long long)gridDim.x * blockDim.x)
        A[(long long)j * lda + i] = 0.0;
}

void zero_small_columns(Context* ctx, double* A, long long m, long long n, long long lda,
                        const double* vals, double rel) {
    if (m <= 0 || n <= 0) return;
    dim3 grid((unsigned)std::min<long long>((m + 255) / 256, 64), (unsigned)n);
    zero_small_columns_kernel<<<grid, 256, 0, ctx->stream>>>(A, m, lda, vals, rel);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

__global__ void sqrt_scalar_kernel(const double* in, double* out) { *out = sqrt(fmax(*in, 0.0)); }

void sqrt_inplace(Context* ctx, const double* dev_in, double* dev_out) {
    sqrt_scalar_kernel<<<1, 1, 0, ctx->stream>>>(dev_in, dev_out);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void select_copy(Context* ctx, double* dst, const double* a, const double* b, long long n,
                 const double* eps_a, const double* eps_b, double* eps_out) {
    select_copy_kernel<<<nblk(n), 256, 0, ctx->stream>>>(dst, a, b, n, eps_a, eps_b, eps_out);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

}  // namespace tnr

// ---- helpers of the block subspace eigensolver (tensor_ops.cu: eigh_topk) ----
namespace tnr {
namespace {

__global__ void fill_random_kernel(double* x, long long n, unsigned long long seed) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    // splitmix64 hash -> uniform(-1, 1); deterministic for a given (seed, i)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    x[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void axpy_kernel(double* y, const double* x, double alpha, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] += alpha * x[i];
}

// two-pass deterministic sum of squares
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const double* __restrict__ x,
                                                            long long n, double* partial) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        double v = x[i];
        acc += v * v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) sum_final_kernel(const double* partial, int n, double* out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

}  // namespace

void fill_random(Context* ctx, double* x, long long n, unsigned long long seed) {
    if (n <= 0) return;
    fill_random_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, n, seed);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void axpy(Context* ctx, double* y, const double* x, double alpha, long long n) {
    if (n <= 0) return;
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(y, x, alpha, n);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

void sum_squares(Context* ctx, const double* x, long long n, double* dev_out) {
    int blocks = (int)std::min<long long>(1024, (n + 255) / 256);
    if (blocks < 1) blocks = 1;
    double* partial = dalloc(ctx, blocks);
    sumsq_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(x, n, partial);
    sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, blocks, dev_out);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches += 2;
    dfree(ctx, partial);
}

}  // namespace tnr
