// FP64 tensor-core GEMM for sm_100a (DMMA.8x8x4 via mma.sync.m8n8k4.f64).
//
// Replaces the BLAS dgemm calls TensorOperations/TensorKit issue for every
// `@tensor` contraction in the reference's step! bodies (e.g.
// /root/reference/src/schemes/hotrg3d.jl:116-120, trg.jl:42).
//
// Design: CTA tile BM x BN x 16, 256 threads, multi-stage cp.async pipeline into
// padded shared memory (conflict-free 64-bit fragment loads for both K-contiguous
// and M/N-contiguous operands, so all four transpose combinations run without a
// separate transpose pass), accumulators in registers, epilogue staged through
// shared memory for coalesced 16-byte stores.  Column major, two-level batch
// strides, optional split-K through a workspace with a deterministic reduction.
#include "common.cuh"

namespace tnr {

namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int PADK = 4;  // row stride (BK+4) == 4 mod 16 doubles -> conflict free
constexpr int PADM = 4;  // row stride (BM+4) == 4 mod 16 doubles

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    long long lda, ldb, ldc;
    int M, N, K;
    double alpha, beta;
    int nb1;
    long long sA1, sA2, sB1, sB2, sC1, sC2;
    int tiles_m, tiles_n;
    int k_per_split;   // multiple of BK; == K rounded up when no split
    int splits;
    double* partial;   // split-K workspace [splits][M*N] (compact ld = M) or null
    int a16, b16, c16; // 16-byte alignment flags
};

// Per-thread loader state for one operand tile (BR rows x BK).  Each thread owns
// PER_THREAD 16-byte chunks whose row / k position inside the tile never changes; only the
// k-tile base advances.  KC: K is the contiguous dimension in global memory.
template <int BR, bool KC>
struct TileLoader {
    static constexpr int CHUNKS = BR * BK / 2;
    static constexpr int PER_THREAD = (CHUNKS + NTHREADS - 1) / NTHREADS;
    const double* g[PER_THREAD];   // global address of the chunk in k-tile 0 (row/k offset applied)
    int soff[PER_THREAD];          // shared offset (doubles) inside a stage
    int kin[PER_THREAD];           // k offset of the chunk inside the tile
    int rvalid[PER_THREAD];        // KC: 2 if row valid else 0 ; !KC: valid elements along rows (0..2)
    const double* base;
    long long kstride;             // global advance per k-tile (doubles)
    bool al16;

    __device__ __forceinline__ void init(const double* gbase, long long ld, int r0, int kbeg, int R,
                                         bool aligned) {
        base = gbase;
        al16 = aligned;
        kstride = KC ? (long long)BK : (long long)BK * ld;
#pragma unroll
        for (int it = 0; it < PER_THREAD; ++it) {
            int c = threadIdx.x + it * NTHREADS;
            int r, k;
            if (KC) {
                r = c / (BK / 2);
                k = (c % (BK / 2)) * 2;
                soff[it] = r * (BK + PADK) + k;
            } else {
                k = c / (BR / 2);
                r = (c % (BR / 2)) * 2;
                soff[it] = k * (BR + PADM) + r;
            }
            if (CHUNKS % NTHREADS != 0 && c >= CHUNKS) {  // thread has no chunk in this slot
                rvalid[it] = -1;
                g[it] = gbase;
                kin[it] = 0;
                continue;
            }
            int gr = r0 + r;
            kin[it] = k;
            if (KC) {
                rvalid[it] = (gr < R) ? 2 : 0;
                g[it] = gbase + (long long)gr * ld + kbeg + k;
            } else {
                rvalid[it] = max(0, min(2, R - gr));
                g[it] = gbase + (long long)(kbeg + k) * ld + gr;
            }
        }
    }

    // issue chunk slot `it` of k-tile `kt` (tile-relative), krem = kend - k0 of this tile
    __device__ __forceinline__ void issue(double* stage, int it, int kt, int krem) const {
        if (rvalid[it] < 0) return;
        int valid;
        if (KC) valid = min(rvalid[it], max(0, krem - kin[it]));
        else valid = (kin[it] < krem) ? rvalid[it] : 0;
        const double* gp = valid ? g[it] + kt * kstride : base;
        double* sp = stage + soff[it];
        if (al16) {
            cp_async16(sp, gp, valid * 8);
        } else {
            cp_async8(sp, gp, valid >= 1 ? 8 : 0);
            cp_async8(sp + 1, valid >= 2 ? gp + 1 : base, valid >= 2 ? 8 : 0);
        }
    }
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool KA, bool KB, int STAGES>
__device__ __forceinline__ void gemm_body(const GemmParams& p, const int tile_id, const int split,
                                          const int b) {
    constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
    constexpr int MI = WTM / 8, NJ = WTN / 8;
    constexpr int A_STAGE = KA ? BM * (BK + PADK) : BK * (BM + PADM);
    constexpr int B_STAGE = KB ? BN * (BK + PADK) : BK * (BN + PADM);
    constexpr int STAGE = A_STAGE + B_STAGE;
    static_assert(WARPS_M * WARPS_N * 32 == NTHREADS, "warp layout");

    extern __shared__ __align__(16) double smem[];

    // ---- tile coordinates (grouped ordering for L2 reuse) ----
    const int GROUP = 8;
    int t = tile_id;
    int per_group = GROUP * p.tiles_n;
    int group_id = t / per_group;
    int first_m = group_id * GROUP;
    int gsize = min(p.tiles_m - first_m, GROUP);
    int tm = first_m + (t % per_group) % gsize;
    int tn = (t % per_group) / gsize;
    const int m0 = tm * BM, n0 = tn * BN;

    const int kbeg = split * p.k_per_split;
    const int kend = min(p.K, kbeg + p.k_per_split);
    const int KT = (kend - kbeg + BK - 1) / BK;

    const int b1 = b % p.nb1, b2 = b / p.nb1;
    const double* __restrict__ A = p.A + b1 * p.sA1 + b2 * p.sA2;
    const double* __restrict__ B = p.B + b1 * p.sB1 + b2 * p.sB2;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lc = lane & 3;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    TileLoader<BM, KA> la;
    TileLoader<BN, KB> lb;
    la.init(A, p.lda, m0, kbeg, p.M, p.a16);
    lb.init(B, p.ldb, n0, kbeg, p.N, p.b16);
    constexpr int PTA = TileLoader<BM, KA>::PER_THREAD, PTB = TileLoader<BN, KB>::PER_THREAD;

    // part q (0..3) of the loads of k-tile kt: interleaved with the four k4-steps of the
    // tile being computed so the DMMA pipe never waits behind a block of address arithmetic
    auto issue_part = [&](int kt, int q) {
        if (kt < KT) {
            double* sa = smem + (kt % STAGES) * STAGE;
            double* sb = sa + A_STAGE;
            int krem = (kend - kbeg) - kt * BK;
#pragma unroll
            for (int it = 0; it < PTA; ++it)
                if (it % 4 == q) la.issue(sa, it, kt, krem);
#pragma unroll
            for (int it = 0; it < PTB; ++it)
                if (it % 4 == q) lb.issue(sb, it, kt, krem);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
#pragma unroll
        for (int q = 0; q < 4; ++q) issue_part(s, q);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const double* sa = smem + (kt % STAGES) * STAGE;
        const double* sb = sa + A_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[MI], bf[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                int r = wm * WTM + 8 * i + lr, k = kk * 4 + lc;
                af[i] = KA ? sa[r * (BK + PADK) + k] : sa[k * (BM + PADM) + r];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int c = wn * WTN + 8 * j + lr, k = kk * 4 + lc;
                bf[j] = KB ? sb[c * (BK + PADK) + k] : sb[k * (BN + PADM) + c];
            }
            issue_part(kt + STAGES - 1, kk);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- epilogue: registers -> smem [col][row] -> coalesced global stores ----
    constexpr int CS = BM + 2;  // == 2 mod 8 -> conflict free fragment scatter
    static_assert(BN * CS <= STAGES * STAGE, "epilogue staging must fit in pipeline smem");
    double* cs = smem;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            int r = wm * WTM + 8 * i + lr;
            int c = wn * WTN + 8 * j + 2 * lc;
            cs[c * CS + r] = acc[i][j][0];
            cs[(c + 1) * CS + r] = acc[i][j][1];
        }
    __syncthreads();

    double* Cout;
    long long ldc;
    double alpha = p.alpha, beta = p.beta;
    bool al16;
    if (p.partial) {
        Cout = p.partial + (long long)split * p.M * p.N;
        ldc = p.M;
        alpha = 1.0;
        beta = 0.0;
        al16 = (p.M % 2 == 0);
    } else {
        Cout = p.C + b1 * p.sC1 + b2 * p.sC2;
        ldc = p.ldc;
        al16 = p.c16;
    }
    constexpr int RC = BM / 2;  // 16-byte chunks per column
    for (int idx = threadIdx.x; idx < BN * RC; idx += NTHREADS) {
        int c = idx / RC, r = (idx % RC) * 2;
        int gr = m0 + r, gc = n0 + c;
        if (gc >= p.N || gr >= p.M) continue;
        double v0 = alpha * cs[c * CS + r];
        double v1 = alpha * cs[c * CS + r + 1];
        double* gp = Cout + (long long)gc * ldc + gr;
        bool two = (gr + 1 < p.M);
        if (beta != 0.0) {
            v0 += beta * gp[0];
            if (two) v1 += beta * gp[1];
        }
        if (two && al16) {
            *reinterpret_cast<double2*>(gp) = make_double2(v0, v1);
        } else {
            gp[0] = v0;
            if (two) gp[1] = v1;
        }
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool KA, bool KB, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_dmma_kernel(const GemmParams p) {
    gemm_body<BM, BN, WARPS_M, WARPS_N, KA, KB, STAGES>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Grouped launch: one kernel covers `nprob` independent problems (the per-sector blocks of a
// symmetric contraction); CTA -> problem through the tile prefix table.
template <int BM, int BN, int WARPS_M, int WARPS_N, bool KA, bool KB, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_dmma_grouped_kernel(const GemmParams* __restrict__ table, const int* __restrict__ tile_start,
                         int nprob) {
    __shared__ GemmParams sp;
    __shared__ int s_first;
    if (threadIdx.x == 0) {
        int g = 0;
        while (g + 1 < nprob && tile_start[g + 1] <= (int)blockIdx.x) ++g;
        sp = table[g];
        s_first = tile_start[g];
    }
    __syncthreads();
    gemm_body<BM, BN, WARPS_M, WARPS_N, KA, KB, STAGES>(sp, blockIdx.x - s_first, 0, 0);
}

__global__ void splitk_reduce_kernel(const double* __restrict__ partial, int splits, long long mn,
                                     int M, double* __restrict__ C, long long ldc, double alpha,
                                     double beta) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= mn) return;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += partial[k * mn + i];
    long long r = i % M, c = i / M;
    double* gp = C + c * ldc + r;
    double v = alpha * s;
    if (beta != 0.0) v += beta * *gp;
    *gp = v;
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool KA, bool KB>
void launch_cfg(Context* ctx, GemmParams& p, int nbatch) {
    constexpr int STAGES = 4;
    constexpr int A_STAGE = KA ? BM * (BK + PADK) : BK * (BM + PADM);
    constexpr int B_STAGE = KB ? BN * (BK + PADK) : BK * (BN + PADM);
    constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    auto kern = gemm_dmma_kernel<BM, BN, WARPS_M, WARPS_N, KA, KB, STAGES>;
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        configured = true;
    }
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    long long tiles = (long long)p.tiles_m * p.tiles_n;
    // split-K when a single (unbatched) problem cannot fill the GPU
    p.splits = 1;
    p.k_per_split = ((p.K + BK - 1) / BK) * BK;
    p.partial = nullptr;
    if (nbatch == 1 && tiles < ctx->num_sms && p.K >= 8 * BK * 4) {
        int want = (int)((2LL * ctx->num_sms + tiles - 1) / tiles);
        int kt = (p.K + BK - 1) / BK;
        int splits = std::min(want, std::max(1, kt / 8));
        if (splits > 1) {
            int kt_per = (kt + splits - 1) / splits;
            splits = (kt + kt_per - 1) / kt_per;
            p.splits = splits;
            p.k_per_split = kt_per * BK;
        }
    }
    double* ws = nullptr;
    if (p.splits > 1) {
        ws = dalloc(ctx, (size_t)p.splits * p.M * p.N);
        p.partial = ws;
    }
    dim3 grid((unsigned)tiles, (unsigned)p.splits, (unsigned)nbatch);
    kern<<<grid, NTHREADS, SMEM, ctx->stream>>>(p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
    ctx->ctr.gemm_launches++;
    if (p.splits > 1) {
        long long mn = (long long)p.M * p.N;
        splitk_reduce_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, ctx->stream>>>(
            ws, p.splits, mn, p.M, p.C, p.ldc, p.alpha, p.beta);
        TNR_CUDA(cudaGetLastError());
        ctx->ctr.launches++;
        dfree(ctx, ws);
    }
}

template <bool KA, bool KB>
void launch_layout(Context* ctx, GemmParams& p, int nbatch) {
    // narrow-N problems (projector applications, N = chi) use the 256x32 tile
    if (p.N <= 48)
        launch_cfg<256, 32, 8, 1, KA, KB>(ctx, p, nbatch);
    else
        launch_cfg<128, 128, 2, 4, KA, KB>(ctx, p, nbatch);
}

}  // namespace

void gemm(Context* ctx, char transa, char transb, int m, int n, int k, double alpha,
          const double* A, long long lda, const double* B, long long ldb, double beta, double* C,
          long long ldc, const GemmBatch& bt) {
    if (m <= 0 || n <= 0) return;
    TNR_CHECK(k > 0, "gemm: k must be positive");
    bool ta = (transa == 'T' || transa == 't'), tb = (transb == 'T' || transb == 't');
    GemmParams p;
    p.A = A; p.B = B; p.C = C;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.M = m; p.N = n; p.K = k;
    p.alpha = alpha; p.beta = beta;
    p.nb1 = bt.nb1;
    p.sA1 = bt.sA1; p.sA2 = bt.sA2; p.sB1 = bt.sB1; p.sB2 = bt.sB2; p.sC1 = bt.sC1; p.sC2 = bt.sC2;
    auto even = [](long long x) { return (x & 1LL) == 0; };
    p.a16 = ((uintptr_t)A % 16 == 0) && even(lda) && even(bt.sA1) && even(bt.sA2);
    p.b16 = ((uintptr_t)B % 16 == 0) && even(ldb) && even(bt.sB1) && even(bt.sB2);
    p.c16 = ((uintptr_t)C % 16 == 0) && even(ldc) && even(bt.sC1) && even(bt.sC2);
    int nbatch = bt.nb1 * bt.nb2;
    TNR_CHECK(nbatch >= 1 && nbatch <= 65535 * 1, "gemm: batch count out of range (<=65535)");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool timed = ctx->time_gemm && (2.0 * m * n * (double)k * nbatch > 1e11);
    if (timed) {
        TNR_CUDA(cudaEventCreate(&e0));
        TNR_CUDA(cudaEventCreate(&e1));
        TNR_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    // KA: A is K-contiguous (transa == 'T');  KB: B is K-contiguous (transb == 'N')
    if (ta && !tb && nbatch == 1 &&
        gemm_tma_tn(ctx, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)) {
        // handled by the TMA kernel
    } else if (ta && !tb) launch_layout<true, true>(ctx, p, nbatch);
    else if (ta && tb) launch_layout<true, false>(ctx, p, nbatch);
    else if (!ta && !tb) launch_layout<false, true>(ctx, p, nbatch);
    else launch_layout<false, false>(ctx, p, nbatch);
    if (timed) {
        TNR_CUDA(cudaEventRecord(e1, ctx->stream));
        ctx->gemm_events.emplace_back(e0, e1);
        ctx->timed_flops += 2.0 * m * n * (double)k * nbatch;
    }
    ctx->ctr.gemm_flops += 2.0 * m * n * (double)k * nbatch;
}

}  // namespace tnr

// ---------------------------------------------------------------------------
// grouped GEMM: all per-sector blocks of a symmetric (Z2/ZN/U1 block-sparse) contraction in
// ONE launch.  Problems share op(A)/op(B) and alpha/beta; sizes and pointers are per problem.
// ---------------------------------------------------------------------------
namespace tnr {
namespace {

template <int BM, int BN, int WARPS_M, int WARPS_N, bool KA, bool KB>
void launch_grouped_cfg(Context* ctx, std::vector<GemmParams>& tab) {
    constexpr int STAGES = 4;
    constexpr int A_STAGE = KA ? BM * (BK + PADK) : BK * (BM + PADM);
    constexpr int B_STAGE = KB ? BN * (BK + PADK) : BK * (BN + PADM);
    constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    auto kern = gemm_dmma_grouped_kernel<BM, BN, WARPS_M, WARPS_N, KA, KB, STAGES>;
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        configured = true;
    }
    std::vector<int> start(tab.size() + 1, 0);
    for (size_t g = 0; g < tab.size(); ++g) {
        GemmParams& p = tab[g];
        p.tiles_m = (p.M + BM - 1) / BM;
        p.tiles_n = (p.N + BN - 1) / BN;
        p.splits = 1;
        p.k_per_split = ((p.K + BK - 1) / BK) * BK;
        p.partial = nullptr;
        start[g + 1] = start[g] + p.tiles_m * p.tiles_n;
    }
    GemmParams* d_tab = nullptr;
    int* d_start = nullptr;
    TNR_CUDA(cudaMallocAsync((void**)&d_tab, tab.size() * sizeof(GemmParams), ctx->stream));
    TNR_CUDA(cudaMallocAsync((void**)&d_start, start.size() * sizeof(int), ctx->stream));
    TNR_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(GemmParams),
                             cudaMemcpyHostToDevice, ctx->stream));
    TNR_CUDA(cudaMemcpyAsync(d_start, start.data(), start.size() * sizeof(int),
                             cudaMemcpyHostToDevice, ctx->stream));
    // pageable sources: the copies above are complete w.r.t. the host buffers on return
    kern<<<(unsigned)start.back(), NTHREADS, SMEM, ctx->stream>>>(d_tab, d_start, (int)tab.size());
    TNR_CUDA(cudaGetLastError());
    TNR_CUDA(cudaFreeAsync(d_tab, ctx->stream));
    TNR_CUDA(cudaFreeAsync(d_start, ctx->stream));
    ctx->ctr.launches++;
    ctx->ctr.gemm_launches++;
    ctx->ctr.grouped_gemm_launches++;
}

template <bool KA, bool KB>
void launch_grouped_layout(Context* ctx, std::vector<GemmParams>& tab, bool narrow) {
    if (narrow) launch_grouped_cfg<256, 32, 8, 1, KA, KB>(ctx, tab);
    else launch_grouped_cfg<128, 128, 2, 4, KA, KB>(ctx, tab);
}

}  // namespace

void gemm_grouped(Context* ctx, char transa, char transb, const std::vector<GroupedProblem>& probs,
                  double alpha, double beta) {
    bool ta = (transa == 'T' || transa == 't'), tb = (transb == 'T' || transb == 't');
    // both operands K-contiguous: the TMA + mbarrier kernel, when every sector fits it
    if (ta && !tb && gemm_grouped_tma_tn(ctx, probs, alpha, beta)) return;
    std::vector<GemmParams> tab;
    bool narrow = true;
    auto even = [](long long x) { return (x & 1LL) == 0; };
    for (const GroupedProblem& q : probs) {
        if (q.m <= 0 || q.n <= 0) continue;
        TNR_CHECK(q.k > 0, "gemm_grouped: k must be positive");
        GemmParams p{};
        p.A = q.A; p.B = q.B; p.C = q.C;
        p.lda = q.lda; p.ldb = q.ldb; p.ldc = q.ldc;
        p.M = q.m; p.N = q.n; p.K = q.k;
        p.alpha = alpha; p.beta = beta;
        p.nb1 = 1;
        p.a16 = ((uintptr_t)q.A % 16 == 0) && even(q.lda);
        p.b16 = ((uintptr_t)q.B % 16 == 0) && even(q.ldb);
        p.c16 = ((uintptr_t)q.C % 16 == 0) && even(q.ldc);
        narrow = narrow && (q.n <= 48);
        tab.push_back(p);
        ctx->ctr.gemm_flops += 2.0 * q.m * q.n * (double)q.k;
    }
    if (tab.empty()) return;
    if (ta && !tb) launch_grouped_layout<true, true>(ctx, tab, narrow);
    else if (ta && tb) launch_grouped_layout<true, false>(ctx, tab, narrow);
    else if (!ta && !tb) launch_grouped_layout<false, true>(ctx, tab, narrow);
    else launch_grouped_layout<false, false>(ctx, tab, narrow);
}

}  // namespace tnr
