// Warp-specialised FP64 tensor-core GEMM for the dominant contraction
// (C = alpha * A^T B + beta * C with both operands K-contiguous: the (f,d)-chunk GEMM of
// HOTRG_3D, /root/reference/src/schemes/hotrg3d.jl:116-120, and every other TN product).
//
//   * one producer warp: a single elected thread drives TMA (cp.async.bulk.tensor.2d, 128-byte
//     swizzle) into a ring of shared-memory stages guarded by full/empty mbarriers;
//   * eight consumer warps: conflict-free 64-bit fragment loads from the swizzled tiles and
//     DMMA.8x8x4 (mma.sync.m8n8k4.f64), accumulators in registers, no CTA-wide barrier in
//     the main loop;
//   * epilogue staged through the (drained) pipeline buffers for coalesced 16-byte stores.
//
// The tensor maps zero-fill out-of-range rows / k, so ragged M, N, K need no special casing.
#include <cuda.h>

#include "common.cuh"

namespace tnr {
namespace {

constexpr int TBM = 128, TBN = 128, TBK = 16;
constexpr int CONSUMER_WARPS = 8;
constexpr int TMA_THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int TSTAGES = 6;
constexpr int A_BYTES = TBM * TBK * 8, B_BYTES = TBN * TBK * 8;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TCS = TBM + 2;  // epilogue staging stride
constexpr size_t TMA_SMEM = 1024 /*align slack*/ + (size_t)TSTAGES * STAGE_BYTES + 256;
static_assert((size_t)TBN * TCS * 8 <= (size_t)TSTAGES * STAGE_BYTES, "epilogue staging");

struct TmaParams {
    double* C;
    long long ldc;
    int M, N, K;
    double alpha, beta;
    int tiles_m, tiles_n;
    int c16;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// fragment row permutation: lane row q (0..7) reads tile row 8*g + rho(q).  With the 128-byte
// swizzle (16-byte chunk index ^= row & 7) this makes the 16 lanes of a half-warp hit 16
// distinct 8-byte banks.
__device__ __forceinline__ int rho(int q) { return 2 * (q & 3) + (q >> 2); }

// One 128 x 128 output tile: `t` is the tile index inside the problem described by p / mapA / mapB.
__device__ __forceinline__ void tma_tile(const CUtensorMap* mapA, const CUtensorMap* mapB,
                                         const TmaParams& p, int t) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned stage ring (128B swizzle atom = 8 rows x 128 B)
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = (unsigned long long*)(smem + (size_t)TSTAGES * STAGE_BYTES);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + TSTAGES);

    const int GROUP = 8;
    int per_group = GROUP * p.tiles_n;
    int group_id = t / per_group;
    int first_m = group_id * GROUP;
    int gsize = min(p.tiles_m - first_m, GROUP);
    int tm = first_m + (t % per_group) % gsize;
    int tn = (t % per_group) / gsize;
    const int m0 = tm * TBM, n0 = tn * TBN;
    const int KT = (p.K + TBK - 1) / TBK;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TSTAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===================== producer warp =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapA));
            asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapB));
            for (int kt = 0; kt < KT; ++kt) {
                int s = kt % TSTAGES;
                int ph = (kt / TSTAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);  // first pass: passes immediately
                unsigned full = full0 + 8 * s;
                mbar_arrive_expect_tx(full, STAGE_BYTES);
                unsigned dstA = smem_u32(smem + (size_t)s * STAGE_BYTES);
                tma_load_2d(dstA, mapA, kt * TBK, m0, full);
                tma_load_2d(dstA + A_BYTES, mapB, kt * TBK, n0, full);
            }
        }
        // the producer warp idles until the consumers are done (no smem reuse hazards: the
        // epilogue staging is only touched after every issued stage has been consumed)
    } else {
        // ===================== consumer warps =====================
        constexpr int WTM = 64, WTN = 32, MI = WTM / 8, NJ = WTN / 8;
        const int wm = warp & 1, wn = warp >> 1;
        const int lr = lane >> 2, lc = lane & 3;
        const int rr = rho(lr);
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // byte offsets of this lane's fragment elements inside a tile, for kk = 0
        // element (row, k): row*128 + (((k>>1) ^ (row&7)) << 4) + (k&1)*8 ; row&7 == rr
        const int a_row_base = (wm * WTM + rr) * 128;
        const int b_row_base = (wn * WTN + rr) * 128;
        const int klo = (lc & 1) * 8;

        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % TSTAGES;
            const int ph = (kt / TSTAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            const unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
            const unsigned char* sb = sa + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < TBK / 4; ++kk) {
                const int chunk = ((2 * kk + (lc >> 1)) ^ rr) << 4;
                double af[MI], bf[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i)
                    af[i] = *reinterpret_cast<const double*>(sa + a_row_base + i * 1024 + chunk + klo);
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    bf[j] = *reinterpret_cast<const double*>(sb + b_row_base + j * 1024 + chunk + klo);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }

        // all consumers done with the ring -> reuse it as C staging [col][row]
        asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_WARPS * 32) : "memory");
        double* cs = reinterpret_cast<double*>(smem);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                int r = wm * WTM + 8 * i + rr;
                int c0 = wn * WTN + 8 * j + rho(2 * lc);
                int c1 = wn * WTN + 8 * j + rho(2 * lc + 1);
                cs[c0 * TCS + r] = acc[i][j][0];
                cs[c1 * TCS + r] = acc[i][j][1];
            }
        asm volatile("bar.sync 1, %0;\n" ::"n"(CONSUMER_WARPS * 32) : "memory");
        const double alpha = p.alpha, beta = p.beta;
        constexpr int RC = TBM / 2;
        for (int idx = threadIdx.x; idx < TBN * RC; idx += CONSUMER_WARPS * 32) {
            int c = idx / RC, r = (idx % RC) * 2;
            int gr = m0 + r, gc = n0 + c;
            if (gc >= p.N || gr >= p.M) continue;
            double v0 = alpha * cs[c * TCS + r];
            double v1 = alpha * cs[c * TCS + r + 1];
            double* gp = p.C + (long long)gc * p.ldc + gr;
            bool two = (gr + 1 < p.M);
            if (beta != 0.0) {
                v0 += beta * gp[0];
                if (two) v1 += beta * gp[1];
            }
            if (two && p.c16) {
                *reinterpret_cast<double2*>(gp) = make_double2(v0, v1);
            } else {
                gp[0] = v0;
                if (two) gp[1] = v1;
            }
        }
    }
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
gemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA,
                     const __grid_constant__ CUtensorMap mapB, const TmaParams p) {
    tma_tile(&mapA, &mapB, p, blockIdx.x);
}

// Grouped launch: the per-sector products of a block-sparse contraction (TensorKit `mul!` over
// `blocks(t)`: src/schemes/trg.jl:42, btrg.jl:86-94 on Z2 / ZN / U1 tensors) in ONE launch of
// the same TMA + mbarrier + DMMA tile.  Every problem has its own pair of tensor maps in a
// device-side table; a CTA finds its problem from the prefix sums of the tile counts.
struct alignas(64) GroupedTmaEntry {
    CUtensorMap mapA, mapB;   // 128 bytes each, 64-byte aligned
    TmaParams p;
    int tile_start;           // first global tile index of this problem
    int pad_;
};

__global__ void __launch_bounds__(TMA_THREADS, 1)
gemm_dmma_tma_grouped_kernel(const GroupedTmaEntry* __restrict__ table, int count) {
    const int t = blockIdx.x;
    int lo = 0, hi = count - 1;             // last entry with tile_start <= t
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].tile_start <= t) lo = mid; else hi = mid - 1;
    }
    const GroupedTmaEntry* e = table + lo;
    tma_tile(&e->mapA, &e->mapB, e->p, t - e->tile_start);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_map(CUtensorMap* map, const double* base, long long rows, long long K, long long ld,
              int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

// Returns false when the problem does not fit the TMA path (caller falls back to the
// cp.async kernel).  A is (K x M, lda) and B is (K x N, ldb), both column major.
bool gemm_tma_tn(Context* ctx, int m, int n, int k, double alpha, const double* A, long long lda,
                 const double* B, long long ldb, double beta, double* C, long long ldc) {
    if (ctx->disable_tma) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return false;
    if (lda * 8 >= (1LL << 40) || ldb * 8 >= (1LL << 40)) return false;
    int tiles_m = (m + TBM - 1) / TBM, tiles_n = (n + TBN - 1) / TBN;
    long long tiles = (long long)tiles_m * tiles_n;
    if (tiles < ctx->num_sms || n <= 48 || k < 4 * TBK) return false;  // split-K / narrow path
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, m, k, lda, TBM) || !make_map(&mapB, B, n, k, ldb, TBN)) return false;
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(gemm_dmma_tma_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        configured = true;
    }
    TmaParams p;
    p.C = C; p.ldc = ldc; p.M = m; p.N = n; p.K = k;
    p.alpha = alpha; p.beta = beta;
    p.tiles_m = tiles_m; p.tiles_n = tiles_n;
    p.c16 = ((uintptr_t)C % 16 == 0) && ((ldc & 1) == 0);
    gemm_dmma_tma_kernel<<<(unsigned)tiles, TMA_THREADS, TMA_SMEM, ctx->stream>>>(mapA, mapB, p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
    ctx->ctr.gemm_launches++;
    ctx->ctr.tma_gemm_launches++;
    return true;
}

// Grouped TN products on the TMA kernel.  Returns false (nothing launched) unless EVERY problem
// fits the TMA path (16-byte aligned K-contiguous operands) and the launch fills the GPU; the
// caller then uses the cp.async grouped kernel.
bool gemm_grouped_tma_tn(Context* ctx, const std::vector<GroupedProblem>& probs, double alpha,
                         double beta) {
    if (ctx->disable_tma || probs.empty()) return false;
    std::vector<GroupedTmaEntry> tab;
    tab.reserve(probs.size());
    long long tiles = 0;
    for (const GroupedProblem& q : probs) {
        if (q.m <= 0 || q.n <= 0) continue;
        if (((uintptr_t)q.A & 15) || ((uintptr_t)q.B & 15) || (q.lda & 1) || (q.ldb & 1)) return false;
        if (q.lda * 8 >= (1LL << 40) || q.ldb * 8 >= (1LL << 40)) return false;
        if (q.k < 4 * TBK) return false;
        GroupedTmaEntry e{};
        if (!make_map(&e.mapA, q.A, q.m, q.k, q.lda, TBM) || !make_map(&e.mapB, q.B, q.n, q.k, q.ldb, TBN))
            return false;
        e.p.C = q.C; e.p.ldc = q.ldc; e.p.M = q.m; e.p.N = q.n; e.p.K = q.k;
        e.p.alpha = alpha; e.p.beta = beta;
        e.p.tiles_m = (q.m + TBM - 1) / TBM; e.p.tiles_n = (q.n + TBN - 1) / TBN;
        e.p.c16 = ((uintptr_t)q.C % 16 == 0) && ((q.ldc & 1) == 0);
        e.tile_start = (int)tiles;
        tiles += (long long)e.p.tiles_m * e.p.tiles_n;
        tab.push_back(e);
    }
    // small launches: the cp.async kernel has split-K and narrow tiles; this one has neither
    if (tab.empty() || tiles < ctx->num_sms || tiles >= (1LL << 31)) return false;
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(gemm_dmma_tma_grouped_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        configured = true;
    }
    const size_t bytes = tab.size() * sizeof(GroupedTmaEntry);
    GroupedTmaEntry* dtab = nullptr;
    TNR_CUDA(cudaMallocAsync((void**)&dtab, bytes, ctx->stream));
    TNR_CHECK(((uintptr_t)dtab & 63) == 0, "grouped TMA table: allocation is not 64-byte aligned");
    // pageable source: the runtime stages the bytes before returning, `tab` may go out of scope
    TNR_CUDA(cudaMemcpyAsync(dtab, tab.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    gemm_dmma_tma_grouped_kernel<<<(unsigned)tiles, TMA_THREADS, TMA_SMEM, ctx->stream>>>(
        dtab, (int)tab.size());
    TNR_CUDA(cudaGetLastError());
    TNR_CUDA(cudaFreeAsync(dtab, ctx->stream));
    for (const GroupedProblem& q : probs)
        if (q.m > 0 && q.n > 0) ctx->ctr.gemm_flops += 2.0 * q.m * q.n * (double)q.k;
    ctx->ctr.launches++;
    ctx->ctr.gemm_launches++;
    ctx->ctr.grouped_gemm_launches++;
    ctx->ctr.tma_gemm_launches++;
    ctx->ctr.tma_grouped_launches++;
    return true;
}

}  // namespace tnr
