// extern "C" boundary of libtnrcuda.so (declared in include/tnrcuda.h).
#include <cstring>

#include "../../include/tnrcuda.h"
#include "schemes.hpp"

using namespace tnr;

struct tnr_context {
    Context c;
};

namespace {

template <class F>
int guard(tnr_context* ctx, F&& f) {
    try {
        f();
        return 0;
    } catch (const Error& e) {
        if (ctx) ctx->c.last_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        if (ctx) ctx->c.last_error = e.what();
        return 3;
    }
}

Dims to_dims(const int64_t* d, int rank) {
    TNR_CHECK(d != nullptr, "null dims");
    Dims r(rank);
    for (int i = 0; i < rank; ++i) {
        TNR_CHECK(d[i] >= 1, "dims must be >= 1");
        r[i] = d[i];
    }
    return r;
}

void copy_out(Context* ctx, const DT& src, double* dst, int64_t* dims_out) {
    TNR_CUDA(cudaMemcpyAsync(dst, src.p, src.size() * sizeof(double), cudaMemcpyDeviceToDevice,
                             ctx->stream));
    if (dims_out)
        for (int i = 0; i < src.rank(); ++i) dims_out[i] = src.d[i];
}

DT in_view(Context* ctx, const double* p, const Dims& d) {
    TNR_CHECK(p != nullptr, "null tensor pointer");
    return DT::view(ctx, const_cast<double*>(p), d);
}

thread_local std::string g_create_error;

}  // namespace

extern "C" {

int tnr_version(void) { return 100; }

int tnr_create(int device, void* stream, tnr_context** out) {
    if (!out) return 1;
    *out = nullptr;
    tnr_context* ctx = new tnr_context();
    int rc = guard(ctx, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Error(2, std::string("libtnrcuda: no CUDA device available (") +
                               cudaGetErrorString(e) + "); there is no CPU fallback");
        TNR_CHECK(device >= 0 && device < ndev, "invalid device ordinal");
        TNR_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        TNR_CUDA(cudaGetDeviceProperties(&prop, device));
        TNR_CHECK(prop.major >= 10, "libtnrcuda is built for sm_100a (Blackwell) only");
        ctx->c.device = device;
        ctx->c.num_sms = prop.multiProcessorCount;
        // NULL selects the legacy default stream (ordered with every blocking stream)
        ctx->c.stream = (cudaStream_t)stream;
        // keep freed blocks cached in the stream-ordered pool (no trim at sync points)
        cudaMemPool_t pool;
        TNR_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        TNR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    });
    if (rc != 0) {
        g_create_error = ctx->c.last_error;
        delete ctx;
        return rc;
    }
    *out = ctx;
    return 0;
}

int tnr_destroy(tnr_context* ctx) {
    if (!ctx) return 1;
    cudaStreamSynchronize(ctx->c.stream);
    for (auto& ev : ctx->c.gemm_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (ctx->c.owns_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
    return 0;
}

const char* tnr_last_error(tnr_context* ctx) {
    return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str();
}

int tnr_synchronize(tnr_context* ctx) {
    if (!ctx) return 1;
    return guard(ctx, [&] { TNR_CUDA(cudaStreamSynchronize(ctx->c.stream)); });
}

int tnr_get_counters(tnr_context* ctx, uint64_t* launches, uint64_t* gemm_launches,
                     double* gemm_flops, double* permute_bytes) {
    if (!ctx) return 1;
    if (launches) *launches = ctx->c.ctr.launches;
    if (gemm_launches) *gemm_launches = ctx->c.ctr.gemm_launches;
    if (gemm_flops) *gemm_flops = ctx->c.ctr.gemm_flops;
    if (permute_bytes) *permute_bytes = ctx->c.ctr.permute_bytes;
    return 0;
}

int tnr_reset_counters(tnr_context* ctx) {
    if (!ctx) return 1;
    ctx->c.ctr = Counters();
    return 0;
}

int tnr_get_tma_launches(tnr_context* ctx, uint64_t* n) {
    if (!ctx || !n) return 1;
    *n = ctx->c.ctr.tma_gemm_launches;
    return 0;
}

int tnr_set_option(tnr_context* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return 1;
    return guard(ctx, [&] {
        if (std::strcmp(key, "disable_tma") == 0) ctx->c.disable_tma = value != 0;
        else if (std::strcmp(key, "ozaki") == 0) {
            TNR_CHECK(value == 0 || (value >= 4 && value <= 10), "ozaki: 0 (off) or 4..10 digit planes");
            ctx->c.ozaki_slices = (int)value;
        }
        else if (std::strcmp(key, "ozaki_crt") == 0) {
            TNR_CHECK(value == 0 || (value >= 14 && value <= 18), "ozaki_crt: 0 (off) or 14..18 moduli");
            ctx->c.ozaki_crt = (int)value;
        }
        else if (std::strcmp(key, "permute_tile") == 0) {
            TNR_CHECK(value == 32 || value == 48 || value == 64 || value == 96,
                      "permute_tile: 32, 48, 64 or 96 (default)");
            ctx->c.permute_tile = (int)value;
        }
        else if (std::strcmp(key, "permute_unroll") == 0) {
            TNR_CHECK(value == 1 || value == 2 || value == 4, "permute_unroll: 1, 2 or 4 (default)");
            ctx->c.permute_unroll = (int)value;
        }
        else if (std::strcmp(key, "permute_bulk") == 0) ctx->c.permute_bulk = value != 0;
        else if (std::strcmp(key, "hotrg3d_pk_budget_mb") == 0) {
            TNR_CHECK(value >= 1, "hotrg3d_pk_budget_mb: megabytes of absorbed operands held at once");
            ctx->c.hotrg3d_pk_budget = (long long)value << 20;
        }
        else if (std::strcmp(key, "permute_dense") == 0) ctx->c.permute_dense = value != 0;
        else if (std::strcmp(key, "permute_tpc") == 0) {
            TNR_CHECK(value >= 1 && value <= 8, "permute_tpc: 1..8");
            ctx->c.permute_tpc = (int)value;
        }
        else if (std::strcmp(key, "permute_chunk_below") == 0) {
            TNR_CHECK(value >= 0 && value <= 65536, "permute_chunk_below: bytes, 0..65536");
            ctx->c.permute_chunk_below = (int)value;
        }
        else if (std::strcmp(key, "disable_subspace") == 0) ctx->c.disable_subspace = value != 0;
        else if (std::strcmp(key, "disable_block_jacobi") == 0) ctx->c.disable_block_jacobi = value != 0;
        else if (std::strcmp(key, "disable_precondition") == 0) ctx->c.disable_precondition = value != 0;
        else if (std::strcmp(key, "disable_qr") == 0) ctx->c.disable_qr = value != 0;
        else if (std::strcmp(key, "disable_cholqr") == 0) ctx->c.disable_cholqr = value != 0;
        else if (std::strcmp(key, "jacobi_max_bc") == 0) {
            TNR_CHECK(value == 4 || value == 8 || value == 16, "jacobi_max_bc: 4, 8 (default) or 16");
            ctx->c.jacobi_max_bc = (int)value;
        }
        else if (std::strcmp(key, "jacobi_max_sweeps") == 0) {
            TNR_CHECK(value >= 1 && value <= 1000, "jacobi_max_sweeps: 1..1000");
            ctx->c.jacobi_max_sweeps = (int)value;
        }
        else if (std::strcmp(key, "disable_persistent_jacobi") == 0) ctx->c.disable_persistent_jacobi = value != 0;
        else throw Error(1, std::string("unknown option: ") + key);
    });
}

int tnr_gemm_timing(tnr_context* ctx, int enable) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
        for (auto& ev : ctx->c.gemm_events) {
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        ctx->c.gemm_events.clear();
        for (auto& ev : ctx->c.phase_events) {
            cudaEventDestroy(ev.e0);
            cudaEventDestroy(ev.e1);
        }
        ctx->c.phase_events.clear();
        ctx->c.timed_flops = 0.0;
        ctx->c.time_gemm = enable != 0;
    });
}

int tnr_gemm_timing_read(tnr_context* ctx, double* ms_total, double* flops_total, int64_t* count) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
        double ms = 0.0;
        for (auto& ev : ctx->c.gemm_events) {
            float t = 0.f;
            TNR_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
            ms += t;
        }
        if (ms_total) *ms_total = ms;
        if (flops_total) *flops_total = ctx->c.timed_flops;
        if (count) *count = (int64_t)ctx->c.gemm_events.size();
    });
}

int tnr_malloc(tnr_context* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return 1;
    return guard(ctx, [&] {
        TNR_CUDA(cudaMallocAsync(dptr, std::max<size_t>(bytes, 8), ctx->c.stream));
    });
}

int tnr_free(tnr_context* ctx, void* dptr) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        if (dptr) TNR_CUDA(cudaFreeAsync(dptr, ctx->c.stream));
    });
}

int tnr_upload(tnr_context* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
    });
}

int tnr_download(tnr_context* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
    });
}

int tnr_gemm(tnr_context* ctx, char transa, char transb, int m, int n, int k, double alpha,
             const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
             int64_t ldc) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(m >= 0 && n >= 0 && k >= 1, "gemm: bad sizes");
        gemm(&ctx->c, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
    });
}

int tnr_gemm_strided_batched(tnr_context* ctx, char transa, char transb, int m, int n, int k,
                             double alpha, const double* A, int64_t lda, int64_t strideA,
                             const double* B, int64_t ldb, int64_t strideB, double beta,
                             double* C, int64_t ldc, int64_t strideC, int batch) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(m >= 0 && n >= 0 && k >= 1 && batch >= 1, "gemm: bad sizes");
        for (int b0 = 0; b0 < batch; b0 += 32768) {
            GemmBatch bt;
            bt.nb1 = std::min(32768, batch - b0);
            bt.sA1 = strideA; bt.sB1 = strideB; bt.sC1 = strideC;
            gemm(&ctx->c, transa, transb, m, n, k, alpha, A + b0 * strideA, lda, B + b0 * strideB,
                 ldb, beta, C + b0 * strideC, ldc, bt);
        }
    });
}

int tnr_gemm_grouped(tnr_context* ctx, char transa, char transb, int count,
                     const tnr_gemm_problem* problems, double alpha, double beta) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(count >= 0 && (count == 0 || problems), "gemm_grouped: bad arguments");
        std::vector<GroupedProblem> v(count);
        for (int g = 0; g < count; ++g) {
            const tnr_gemm_problem& q = problems[g];
            v[g] = GroupedProblem{q.m, q.n, q.k, q.A, q.lda, q.B, q.ldb, q.C, q.ldc};
        }
        gemm_grouped(&ctx->c, transa, transb, v, alpha, beta);
    });
}

int tnr_gemm_ozaki(tnr_context* ctx, int m, int n, int k, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double* C, int64_t ldc) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        TNR_CHECK(c->ozaki_slices > 0 || c->ozaki_crt > 0,
                  "gemm_ozaki: enable with tnr_set_option(ctx, \"ozaki\", 8) or (ctx, \"ozaki_crt\", 16)");
        TNR_CHECK(ozaki_applicable(c, m, n, k),
                  "gemm_ozaki: needs m, n, k >= 512, k % 16 == 0 and slices * k * 4096 < 2^31");
        OzakiOperand a = ozaki_split(c, A, lda, m, k);
        OzakiOperand b = ozaki_split(c, B, ldb, n, k);
        ozaki_multiply(c, a, b, C, ldc);
        ozaki_free(c, a);
        ozaki_free(c, b);
    });
}

int tnr_permute(tnr_context* ctx, const double* src, double* dst, int rank, const int64_t* dims,
                const int* perm) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Dims d = to_dims(dims, rank);
        permute(&ctx->c, src, dst, rank, d.data(), perm);
    });
}

int tnr_contract(tnr_context* ctx, const double* A, int rankA, const int64_t* dimsA,
                 const char* labelsA, const double* B, int rankB, const int64_t* dimsB,
                 const char* labelsB, double* C, const char* labelsC) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(labelsA && labelsB && labelsC, "contract: null labels");
        DT a = in_view(&ctx->c, A, to_dims(dimsA, rankA));
        DT b = in_view(&ctx->c, B, to_dims(dimsB, rankB));
        DT c = contract(a, labelsA, b, labelsB, labelsC);
        copy_out(&ctx->c, c, C, nullptr);
    });
}

int tnr_svd_trunc(tnr_context* ctx, const double* T, int rank, const int64_t* dims, int ncod,
                  int chi, double* U, double* S, double* Vt, int64_t* k_out, double* eps_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(ncod >= 1 && ncod < rank && chi >= 1, "svd_trunc: bad arguments");
        DT t = in_view(&ctx->c, T, to_dims(dims, rank));
        Trunc f = svd_trunc(t, ncod, chi);
        copy_out(&ctx->c, f.U, U, nullptr);
        copy_out(&ctx->c, f.S, S, nullptr);
        copy_out(&ctx->c, f.Vt, Vt, nullptr);
        if (k_out) *k_out = f.S.size();
        double e = 0.0;
        TNR_CUDA(cudaMemcpyAsync(&e, f.eps.p, 8, cudaMemcpyDeviceToHost, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
        if (eps_out) *eps_out = e;
    });
}

int tnr_qr(tnr_context* ctx, const double* A, int64_t m, int64_t n, double* Q, double* R) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(A && m >= n && n >= 1 && (Q || R), "qr: bad arguments");
        qr_thin(&ctx->c, A, m, n, Q, R);
    });
}

int tnr_orth_r(tnr_context* ctx, const double* T, int rank, const int64_t* dims, int ncod,
               double* R) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(ncod >= 1 && ncod < rank, "orth_r: bad arguments");
        DT t = in_view(&ctx->c, T, to_dims(dims, rank));
        DT r = orth_r(t, ncod);
        copy_out(&ctx->c, r, R, nullptr);
    });
}

int tnr_psd_factor(tnr_context* ctx, const double* G, int64_t n, double* L, int64_t* rank_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(G && L && n >= 1, "psd_factor: bad arguments");
        const long long r = psd_factor(&ctx->c, G, n, L);
        if (rank_out) *rank_out = r;
    });
}

int tnr_fill_random(tnr_context* ctx, double* x, int64_t n, uint64_t seed) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(x && n >= 0, "fill_random: bad arguments");
        fill_random(&ctx->c, x, n, seed);
    });
}

int tnr_orthonormalize(tnr_context* ctx, double* A, int64_t m, int64_t n, int* refused_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(A && refused_out && m >= n && n >= 1, "orthonormalize: bad arguments");
        *refused_out = cholqr2(&ctx->c, A, m, n) ? 0 : 1;
    });
}

int tnr_eigh_trunc(tnr_context* ctx, const double* MM, int64_t n, int chi, double* W, double* V,
                   int64_t* k_out, double* eps_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(n >= 1 && chi >= 1, "eigh_trunc: bad arguments");
        DT m = in_view(&ctx->c, MM, {n, n});
        Trunc f = eigh_trunc(clone(m), 1, chi);
        copy_out(&ctx->c, f.S, W, nullptr);
        copy_out(&ctx->c, f.U, V, nullptr);
        if (k_out) *k_out = f.S.size();
        double e = 0.0;
        TNR_CUDA(cudaMemcpyAsync(&e, f.eps.p, 8, cudaMemcpyDeviceToHost, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
        if (eps_out) *eps_out = e;
    });
}

int tnr_step_out_dims(int scheme, const int64_t* dims, int chi, int64_t* o) {
    if (!dims || !o || chi < 1) return 1;
    auto mn = [&](int64_t a) { return std::min<int64_t>(chi, a); };
    switch (scheme) {
        case TNR_TRG:
        case TNR_BTRG: {
            // bonds: first SVD (1 2 | 3 4) -> k1, second (2 4 | 1 3) resp. (3 1 | 4 2) -> k2
            int64_t k1 = mn(std::min(dims[0] * dims[1], dims[2] * dims[3]));
            int64_t k2 = mn(std::min(dims[1] * dims[3], dims[0] * dims[2]));
            if (scheme == TNR_TRG) { o[0] = k1; o[1] = k2; o[2] = k2; o[3] = k1; }
            else { o[0] = k2; o[1] = k1; o[2] = k1; o[3] = k2; }
            return 0;
        }
        case TNR_HOTRG: {
            int64_t kx = mn(dims[0] * dims[0]);
            int64_t ky = mn(dims[1] * dims[1]);
            o[0] = kx; o[1] = ky; o[2] = ky; o[3] = kx;
            return 0;
        }
        case TNR_ATRG: {
            int64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
            auto half = [&](int64_t* x) {
                int64_t k1 = mn(std::min(x[0] * x[2], x[1] * x[3]));
                int64_t k2 = mn(std::min(x[0] * k1, k1 * x[3]));
                int64_t k3 = mn(std::min(k2 * x[1], x[2] * k2));
                x[0] = k3; x[3] = k3;  // T = (k3, d1, d2, k3)
            };
            half(d);
            int64_t t[4] = {d[1], d[3], d[0], d[2]};  // ((2,4),(1,3))
            half(t);
            o[0] = t[2]; o[1] = t[0]; o[2] = t[3]; o[3] = t[1];  // ((3,1),(4,2))
            return 0;
        }
        case TNR_HOTRG_3D:
        case TNR_ATRG_3D: {
            int64_t d[6];
            for (int i = 0; i < 6; ++i) d[i] = dims[i];
            const int ph[6] = {5, 3, 1, 2, 0, 4};  // ((6,4),(2,3,1,5))
            const int pa[6] = {3, 5, 1, 4, 0, 2};  // ((4,6),(2,5,1,3))
            const int* p = (scheme == TNR_HOTRG_3D) ? ph : pa;
            for (int it = 0; it < 3; ++it) {
                int64_t nx = mn(d[3] * d[3]), ny = mn(d[2] * d[2]);
                int64_t t[6] = {d[0], d[1], ny, nx, ny, nx};
                for (int i = 0; i < 6; ++i) d[i] = t[p[i]];
            }
            for (int i = 0; i < 6; ++i) o[i] = d[i];
            return 0;
        }
        default:
            return 1;
    }
}

int tnr_trg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                 int64_t* dims_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 4));
        DT r = trg_step(&ctx->c, t, chi);
        copy_out(&ctx->c, r, Tout, dims_out);
    });
}

int tnr_btrg_step(tnr_context* ctx, const double* T, const int64_t* dims, const double* S1,
                  const double* S2, double k, int chi, double* Tout, int64_t* dims_out,
                  double* S1out, double* S2out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        Dims d = to_dims(dims, 4);
        DT t = clone(in_view(c, T, d));
        DT s1 = clone(in_view(c, S1, {d[1]}));
        DT s2 = clone(in_view(c, S2, {d[0]}));
        btrg_step(c, t, s1, s2, k, chi);
        copy_out(c, t, Tout, dims_out);
        copy_out(c, s1, S1out, nullptr);
        copy_out(c, s2, S2out, nullptr);
    });
}

int tnr_hotrg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                   int64_t* dims_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 4));
        DT r = hotrg_step(&ctx->c, t, chi);
        copy_out(&ctx->c, r, Tout, dims_out);
    });
}

int tnr_atrg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                  int64_t* dims_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 4));
        DT r = atrg_step(&ctx->c, t, chi);
        copy_out(&ctx->c, r, Tout, dims_out);
    });
}

int tnr_hotrg3d_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                     double* Tout, int64_t* dims_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 6));
        DT r = hotrg3d_step(&ctx->c, t, chi);
        copy_out(&ctx->c, r, Tout, dims_out);
    });
}

int tnr_hotrg3d_substep(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                        double* Tout, int64_t* dims_out, int64_t f_begin, int64_t f_end) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        DT t = in_view(c, T, to_dims(dims, 6));
        Dims od = hotrg3d_substep_dims(t.d, chi);
        TNR_CHECK(0 <= f_begin && f_begin <= f_end && f_end <= od[5], "substep: bad slice range");
        DT out = DT::view(c, Tout, od);
        hotrg3d_substep(c, t, chi, out, f_begin, f_end, nullptr, 0);
        if (dims_out)
            for (int i = 0; i < 6; ++i) dims_out[i] = od[i];
    });
}

int tnr_hotrg3d_substep_peers(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                              double* const* Tout_peers, int npeers, int self, int64_t* dims_out,
                              int64_t f_begin, int64_t f_end) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        TNR_CHECK(Tout_peers && npeers >= 1 && npeers <= 16 && self >= 0 && self < npeers,
                  "substep_peers: bad peer table");
        DT t = in_view(c, T, to_dims(dims, 6));
        Dims od = hotrg3d_substep_dims(t.d, chi);
        TNR_CHECK(0 <= f_begin && f_begin <= f_end && f_end <= od[5], "substep: bad slice range");
        DT out = DT::view(c, Tout_peers[self], od);
        hotrg3d_substep(c, t, chi, out, f_begin, f_end, Tout_peers, npeers);
        if (dims_out)
            for (int i = 0; i < 6; ++i) dims_out[i] = od[i];
    });
}

int tnr_hotrg3d_proj_half(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                          int which, double* out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        TNR_CHECK(out && which >= 0 && which < 4, "proj_half: bad arguments");
        DT t = in_view(c, T, to_dims(dims, 6));
        PhaseScope ph(c, "hotrg3d.projectors");
        Trunc h = hotrg3d_proj_half(c, t, which, chi);
        const long long nk = h.U.size();
        TNR_CUDA(cudaMemcpyAsync(out, h.U.p, nk * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        TNR_CUDA(cudaMemcpyAsync(out + nk, h.eps.p, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    });
}

int tnr_hotrg3d_contract(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                         const double* halves, double* const* Tout_peers, int npeers, int self,
                         int64_t* dims_out, int64_t f_begin, int64_t f_end) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Context* c = &ctx->c;
        TNR_CHECK(halves && Tout_peers && npeers >= 1 && npeers <= 16 && self >= 0 && self < npeers,
                  "hotrg3d_contract: bad arguments");
        DT t = in_view(c, T, to_dims(dims, 6));
        Dims od = hotrg3d_substep_dims(t.d, chi);
        TNR_CHECK(0 <= f_begin && f_begin <= f_end && f_end <= od[5], "contract: bad slice range");
        const long long Dy = t.d[2], Dx = t.d[3], ny = od[2], nx = od[3];
        const long long sx = Dx * Dx * nx + 1, sy = Dy * Dy * ny + 1;
        const double* hx0 = halves;
        const double* hx1 = halves + sx;
        const double* hy0 = halves + 2 * sx;
        const double* hy1 = halves + 2 * sx + sy;
        DT Ux = hotrg3d_pick(c, DT::view(c, const_cast<double*>(hx0), {Dx, Dx, nx}), hx0 + sx - 1,
                             DT::view(c, const_cast<double*>(hx1), {Dx, Dx, nx}), hx1 + sx - 1);
        DT Uy = hotrg3d_pick(c, DT::view(c, const_cast<double*>(hy0), {Dy, Dy, ny}), hy0 + sy - 1,
                             DT::view(c, const_cast<double*>(hy1), {Dy, Dy, ny}), hy1 + sy - 1);
        DT out = DT::view(c, Tout_peers[self], od);
        hotrg3d_contract(c, t, Ux, Uy, out, f_begin, f_end, npeers > 1 ? Tout_peers : nullptr,
                         npeers > 1 ? npeers : 0);
        if (dims_out)
            for (int i = 0; i < 6; ++i) dims_out[i] = od[i];
    });
}

int tnr_atrg3d_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                    int64_t* dims_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 6));
        DT r = atrg3d_step(&ctx->c, t, chi);
        copy_out(&ctx->c, r, Tout, dims_out);
    });
}

int tnr_finalize_2d(tnr_context* ctx, double* T, const int64_t* dims, double* norm_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 4));
        double n = finalize_2d(&ctx->c, t);
        if (norm_out) *norm_out = n;
    });
}

int tnr_finalize_btrg(tnr_context* ctx, double* T, const int64_t* dims, const double* S1,
                      const double* S2, double* norm_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        Dims d = to_dims(dims, 4);
        DT t = in_view(&ctx->c, T, d);
        DT s1 = in_view(&ctx->c, S1, {d[1]});
        DT s2 = in_view(&ctx->c, S2, {d[0]});
        double n = finalize_btrg(&ctx->c, t, s1, s2);
        if (norm_out) *norm_out = n;
    });
}

int tnr_finalize_3d(tnr_context* ctx, double* T, const int64_t* dims, double* norm_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        DT t = in_view(&ctx->c, T, to_dims(dims, 6));
        double n = finalize_3d(&ctx->c, t);
        if (norm_out) *norm_out = n;
    });
}

}  // extern "C"

// ---- small primitives used by the block-sparse (abelian sector) host layer ----
extern "C" {

int tnr_strided_copy(tnr_context* ctx, const double* src, double* dst, int rank,
                     const int64_t* dims, const int64_t* sstride, const int64_t* dstride) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(rank >= 0 && rank <= 16 && dims && sstride && dstride, "strided_copy: bad args");
        long long d[16], ss[16], ds[16];
        for (int i = 0; i < rank; ++i) { d[i] = dims[i]; ss[i] = sstride[i]; ds[i] = dstride[i]; }
        strided_copy(&ctx->c, src, dst, rank, d, ss, ds);
    });
}

int tnr_strided_sum(tnr_context* ctx, const double* src, int rank, const int64_t* dims,
                    const int64_t* stride, const double* const* weights, double* sum_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(rank >= 1 && rank <= 4 && dims && stride && sum_out, "strided_sum: bad args");
        long long d[4], st[4];
        for (int i = 0; i < rank; ++i) { d[i] = dims[i]; st[i] = stride[i]; }
        DT out(&ctx->c, {1});
        strided_sum_w(&ctx->c, src, rank, d, st, weights, out.p, false);
        TNR_CUDA(cudaMemcpyAsync(sum_out, out.p, 8, cudaMemcpyDeviceToHost, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
    });
}

int tnr_scale(tnr_context* ctx, double* x, int64_t n, double alpha) {
    if (!ctx) return 1;
    return guard(ctx, [&] { scale(&ctx->c, x, n, alpha); });
}

int tnr_diag_scale(tnr_context* ctx, double* A, int64_t m, int64_t n, int64_t lda, const double* s,
                   int rows, int mode, double p) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(mode >= 0 && mode <= 2, "diag_scale: mode must be 0 (s), 1 (sqrt s), 2 (pseudopow)");
        diag_scale(&ctx->c, A, m, n, lda, s, rows != 0, mode, p);
    });
}

int tnr_vec_map(tnr_context* ctx, const double* s, double* out, int64_t n, int mode, double p) {
    if (!ctx) return 1;
    return guard(ctx, [&] { vec_map(&ctx->c, s, out, n, mode, p); });
}

int tnr_topk_select(tnr_context* ctx, const double* vals, int64_t n, int64_t k, int32_t* rank_host,
                    double* eps_out) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(n >= 1 && k >= 0 && rank_host, "topk_select: bad args");
        int* d_rank = nullptr;
        TNR_CUDA(cudaMallocAsync((void**)&d_rank, n * sizeof(int), ctx->c.stream));
        DT eps(&ctx->c, {1});
        rank_select(&ctx->c, vals, n, k, d_rank, eps.p);
        double e = 0.0;
        TNR_CUDA(cudaMemcpyAsync(rank_host, d_rank, n * sizeof(int), cudaMemcpyDeviceToHost,
                                 ctx->c.stream));
        TNR_CUDA(cudaMemcpyAsync(&e, eps.p, 8, cudaMemcpyDeviceToHost, ctx->c.stream));
        TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
        TNR_CUDA(cudaFreeAsync(d_rank, ctx->c.stream));
        if (eps_out) *eps_out = e;
    });
}

}  // extern "C"

extern "C" int tnr_axis_scale(tnr_context* ctx, double* A, int64_t m1, int64_t n, int64_t m2,
                              const double* s, int mode, double p) {
    if (!ctx) return 1;
    return guard(ctx, [&] {
        TNR_CHECK(mode >= 0 && mode <= 2 && m1 >= 1 && n >= 1 && m2 >= 1, "axis_scale: bad args");
        axis_scale(&ctx->c, A, m1, n, m2, s, mode, p);
    });
}

extern "C" int tnr_get_counter(tnr_context* ctx, const char* name, double* value) {
    if (!ctx || !name || !value) return 1;
    return guard(ctx, [&] {
        const Counters& c = ctx->c.ctr;
        std::string n(name);
        if (n.rfind("phase_ms.", 0) == 0) {
            // summed CUDA-event milliseconds of a named phase since tnr_gemm_timing(1)
            TNR_CUDA(cudaStreamSynchronize(ctx->c.stream));
            const std::string want = n.substr(9);
            double ms = 0.0;
            for (auto& ev : ctx->c.phase_events) {
                if (want != ev.name) continue;
                float t = 0.f;
                TNR_CUDA(cudaEventElapsedTime(&t, ev.e0, ev.e1));
                ms += t;
            }
            *value = ms;
            return;
        }
        if (n == "launches") *value = (double)c.launches;
        else if (n == "gemm_launches") *value = (double)c.gemm_launches;
        else if (n == "grouped_gemm_launches") *value = (double)c.grouped_gemm_launches;
        else if (n == "tma_gemm_launches") *value = (double)c.tma_gemm_launches;
        else if (n == "tma_grouped_launches") *value = (double)c.tma_grouped_launches;
        else if (n == "ozaki_launches") *value = (double)c.ozaki_launches;
        else if (n == "ozaki_gemms") *value = (double)c.ozaki_gemms;
        else if (n == "peer_scatter_launches") *value = (double)c.peer_scatter_launches;
        else if (n == "permute_bulk_launches") *value = (double)c.permute_bulk_launches;
        else if (n == "preconditioned_jacobi") *value = (double)c.preconditioned_jacobi;
        else if (n == "subspace_eigh") *value = (double)c.subspace_eigh;
        else if (n == "subspace_svd") *value = (double)c.subspace_svd;
        else if (n == "persistent_jacobi") *value = (double)c.persistent_jacobi;
        else if (n == "qr_factorizations") *value = (double)c.qr_factorizations;
        else if (n == "psd_factorizations") *value = (double)c.psd_factorizations;
        else if (n == "cholqr2") *value = (double)c.cholqr2;
        else if (n == "cholqr2_refused") *value = (double)c.cholqr2_refused;
        else if (n == "jacobi_limit_accepted") *value = (double)c.jacobi_limit_accepted;
        else if (n == "jacobi_not_converged") *value = (double)c.jacobi_not_converged;
        else if (n == "subspace_fallbacks") *value = (double)c.subspace_fallbacks;
        else if (n == "gemm_flops") *value = c.gemm_flops;
        else if (n == "permute_bytes") *value = c.permute_bytes;
        else throw Error(1, "unknown counter: " + n);
    });
}
