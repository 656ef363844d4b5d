// FP64-accurate GEMM on the INT8 tensor cores (Ozaki scheme) -- opt-in engine for the
// dominant TN contraction  C(m x n) = A^T B,  A: K x M, B: K x N, column major
// (the (f,d)-chunk GEMM of HOTRG_3D, /root/reference/src/schemes/hotrg3d.jl:116-120).
//
// The FP64 tensor pipe (DMMA) tops out at ~36 TFLOP/s and the chunk GEMM already runs at 98 % of
// that; the INT8 path of the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators
// in TMEM) delivers ~2.3 POPS with this kernel.  Error-free splitting turns one FP64 product into
// S(S+1)/2 exact INT8 products:
//
//   * every row x of A^T (and of B^T) is scaled by a power of two, |y| = |x| 2^-e < 1, and
//     expanded in signed base-128 digits  y = sum_i q_i 2^(-6-7i),  |q_i| <= 64  (exact in FP64);
//   * digit planes are multiplied exactly in INT8 x INT8 -> INT32 (K * 64^2 * S < 2^31);
//     all pairs with i + j = t share one TMEM accumulator (the planes are simply concatenated
//     along K), so there are S accumulation groups t = 0..S-1 with weight 2^(-12-7t);
//   * groups are summed into C in FP64 from the smallest weight to the largest, scaled by
//     2^(e_row + e_col).
//
// With S = 8 the neglected pairs (i + j >= 8) are below 9 * 2^-56 ~ 1.3e-16 per term relative
// to (row max) x (column max): the truncation error is of the size of DGEMM's own rounding.
//
// Kernel structure: one TMA producer thread (3-D tensor maps: K x rows x plane, 128-byte
// swizzle), one MMA thread issuing tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 256, K = 32)
// into a 128 x 256 INT32 accumulator in TMEM, four epilogue warps (tcgen05.ld 32x32b) that
// convert, scale and accumulate into the FP64 result.
#include <cuda.h>

#include <cmath>

#include "common.cuh"
#include "crt_math.cuh"

namespace tnr {
namespace {

// constants of the CRT variant (one table per launch configuration, set by ozaki_crt_table)
__constant__ CrtTable c_crt;

constexpr int OBM = 128, OBN = 256, OBK = 128, OSTAGES = 4;
constexpr int OA_BYTES = OBM * OBK, OB_BYTES = OBN * OBK, OSTAGE_BYTES = OA_BYTES + OB_BYTES;
constexpr int OTHREADS = 192;
constexpr int OTMEM_COLS = 256;
constexpr size_t OSMEM = 1024 + (size_t)OSTAGES * OSTAGE_BYTES + 128;

__device__ __forceinline__ unsigned o_smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void o_mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void o_mbar_expect_tx(unsigned bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// bounded wait: a protocol error becomes a trap (reported as a CUDA error), never a hang
__device__ __forceinline__ void o_mbar_wait(unsigned bar, int parity) {
    for (long long it = 0; it < (1LL << 30); ++it) {
        unsigned ok;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    asm volatile("trap;\n");
}
__device__ __forceinline__ void o_tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1,
                                              int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// UMMA shared-memory descriptor: K-major, 128-byte swizzle, LBO = 16 B, SBO = 1024 B, version 1
__device__ __forceinline__ uint64_t o_make_desc(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void o_umma_i8(unsigned tmem_c, uint64_t adesc, uint64_t bdesc,
                                          unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void o_umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                 : "memory");
}

struct OzakiParams {
    double* C;
    long long ldc;
    const double* scaleA;  // 2^e per row of A^T
    const double* scaleB;  // 2^e per row of B^T
    int M, N, K;
    int S;         // digit planes; groups t = S-1 .. 0, group t = pairs (i, t - i), weight 2^(-12-7t)
                   // CRT variant: number of moduli; group g = the single pair (g, g)
    unsigned char* R;        // CRT variant: residues [S][N][ldr] of the accumulators, in [0, p_g)
    long long ldr;           // rows of R (>= M)
};

// One CTA per 128 x 256 output tile.  All S accumulation groups of the tile run inside one
// launch: the MMA thread alternates between two 256-column TMEM accumulators, the four epilogue
// warps drain accumulator g (convert, scale, accumulate into FP64 C) while the MMAs of group
// g + 1 are already running.
//
// CRT = true (option "ozaki_crt"): the planes are the symmetric residues of the scaled operands
// modulo p_0..p_{S-1}; group g is the single product (plane g) x (plane g), and the epilogue
// stores the residue of the INT32 accumulator mod p_g as one byte per element
// (ozaki_crt_reconstruct_kernel turns the S residues into the FP64 result afterwards).
template <bool CRT>
__global__ void __launch_bounds__(OTHREADS, 1)
ozaki_tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const OzakiParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = (unsigned long long*)(smem + (size_t)OSTAGES * OSTAGE_BYTES);
    const unsigned full0 = o_smem_u32(bars), empty0 = o_smem_u32(bars + OSTAGES);
    const unsigned tfull0 = o_smem_u32(bars + 2 * OSTAGES);        // [2] accumulator ready
    const unsigned tempty0 = o_smem_u32(bars + 2 * OSTAGES + 2);   // [2] accumulator drained
    unsigned* tmem_ptr = (unsigned*)(bars + 2 * OSTAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grouped tile order: the ~148 concurrently resident CTAs cover a near-square patch of
    // output tiles, so every A / B plane tile is fetched from HBM once and reused from L2
    const int tiles_m = (p.M + OBM - 1) / OBM, tiles_n = (p.N + OBN - 1) / OBN;
    constexpr int GROUP = 12;
    const int per_group = GROUP * tiles_n;
    const int gid = blockIdx.x / per_group;
    const int first_m = gid * GROUP;
    const int gsz = min(tiles_m - first_m, GROUP);
    const int tm = first_m + (blockIdx.x % per_group) % gsz;
    const int tn = (blockIdx.x % per_group) / gsz;
    const int m0 = tm * OBM, n0 = tn * OBN;
    const int KB = (p.K + OBK - 1) / OBK;
    const int S = p.S;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OSTAGES; ++s) {
            o_mbar_init(full0 + 8 * s, 1);
            o_mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            o_mbar_init(tfull0 + 8 * b, 1);
            o_mbar_init(tempty0 + 8 * b, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         o_smem_u32(tmem_ptr)),
                     "n"(2 * OTMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int g = 0; g < S; ++g) {
                const int t = S - 1 - g;
                for (int i = CRT ? g : 0; i <= (CRT ? g : t); ++i) {
                    const int j = CRT ? g : t - i;
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        int s = it % OSTAGES, ph = (it / OSTAGES) & 1;
                        o_mbar_wait(empty0 + 8 * s, ph ^ 1);
                        unsigned full = full0 + 8 * s;
                        o_mbar_expect_tx(full, OSTAGE_BYTES);
                        unsigned dst = o_smem_u32(smem + (size_t)s * OSTAGE_BYTES);
                        o_tma_load_3d(dst, &mapA, kb * OBK, m0, i, full);
                        o_tma_load_3d(dst + OA_BYTES, &mapB, kb * OBK, n0, j, full);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D = S32, A = B = signed int8, both K-major, N = 256, M = 128
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) |
                                   ((unsigned)(OBN >> 3) << 17) | ((unsigned)(OBM >> 4) << 24);
            int it = 0;
            for (int g = 0; g < S; ++g) {
                const int t = S - 1 - g, buf = g & 1;
                o_mbar_wait(tempty0 + 8 * buf, ((g >> 1) & 1) ^ 1);  // accumulator drained
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned acc = tmem_base + (unsigned)(buf * OTMEM_COLS);
                const int nblk = CRT ? KB : KB * (t + 1);
                for (int b = 0; b < nblk; ++b, ++it) {
                    int s = it % OSTAGES, ph = (it / OSTAGES) & 1;
                    o_mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    unsigned a_base = o_smem_u32(smem + (size_t)s * OSTAGE_BYTES);
                    unsigned b_base = a_base + OA_BYTES;
#pragma unroll
                    for (int k = 0; k < OBK / 32; ++k)
                        o_umma_i8(acc, o_make_desc(a_base + k * 32), o_make_desc(b_base + k * 32),
                                  idesc, (b > 0 || k > 0) ? 1u : 0u);
                    o_umma_commit(empty0 + 8 * s);
                }
                o_umma_commit(tfull0 + 8 * buf);
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const int row = m0 + q * 32 + lane;
        const double sa = (row < p.M) ? p.scaleA[row] : 0.0;
        for (int g = 0; g < S; ++g) {
            const int t = S - 1 - g, buf = g & 1;
            const double w = sa * exp2((double)(-12 - 7 * t));
            o_mbar_wait(tfull0 + 8 * buf, (g >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int c0 = 0; c0 < OBN; c0 += 32) {
                unsigned v[32];
                unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) +
                                 (unsigned)(buf * OTMEM_COLS + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                      "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                      "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                      "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                      "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                if constexpr (CRT) {
                    if (row < p.M) {
                        const int pg = c_crt.p[g];
                        const double ig = c_crt.inv_p[g];
                        unsigned char* rp = p.R + ((long long)g * p.N + (n0 + c0)) * p.ldr + row;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + c0 + j < p.N)
                                rp[(long long)j * p.ldr] =
                                    (unsigned char)crt_acc_residue((int)v[j], pg, ig);
                    }
                } else
                if (row < p.M) {
                    // all loads of the old C values are issued before the first store, so the
                    // 32 read-modify-writes overlap instead of paying 32 global latencies in turn
                    double old[32];
                    double* cp = p.C + (long long)(n0 + c0) * p.ldc + row;
                    if (g > 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            old[j] = (n0 + c0 + j < p.N) ? __ldcg(cp + (long long)j * p.ldc) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        int col = n0 + c0 + j;
                        if (col < p.N) {
                            double val = (double)(int)v[j] * (w * p.scaleB[col]);
                            __stcg(cp + (long long)j * p.ldc, (g > 0) ? (old[j] + val) : val);
                        }
                    }
                }
            }
            // accumulator `buf` may be overwritten by group g + 2
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tempty0 + 8 * buf)
                             : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                     "n"(2 * OTMEM_COLS)
                     : "memory");
    }
}

// ---- error-free splitting of the rows of a K-major FP64 matrix into int8 digit planes ----
// X: K x R column major (row r of X^T = contiguous column r, leading dimension ld).
// planes[i][r][k] (int8, k contiguous), scale[r] = 2^e with |x| 2^-e < 1.
__global__ void __launch_bounds__(256) ozaki_split_kernel(const double* __restrict__ X,
                                                          long long ld, int K, int slices,
                                                          int8_t* __restrict__ planes,
                                                          long long plane_stride,
                                                          double* __restrict__ scale) {
    const int r = blockIdx.x;
    const double* x = X + (long long)r * ld;
    __shared__ double red[256];
    double amax = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) amax = fmax(amax, fabs(x[k]));
    red[threadIdx.x] = amax;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    amax = red[0];
    int e = 0;
    if (amax > 0.0 && isfinite(amax)) e = ilogb(amax) + 1;  // amax * 2^-e in [0.5, 1)
    if (threadIdx.x == 0) scale[r] = ldexp(1.0, e);
    int8_t* out = planes + (long long)r * K;
    // 8 consecutive k per thread: one 64-byte read, one packed 8-byte store per digit plane
    // (K is a multiple of 16, rows of the planes are 8-byte aligned)
    for (int k8 = threadIdx.x * 8; k8 < K; k8 += 256 * 8) {
        double rem[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) rem[u] = ldexp(x[k8 + u], 6 - e) * (1.0 / 128.0);
        for (int i = 0; i < slices; ++i) {
            unsigned long long pack = 0ULL;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                double y = rem[u] * 128.0;     // exact
                double q = rint(y);            // |q| <= 64
                rem[u] = y - q;                // |rem| <= 0.5, exact
                pack |= (unsigned long long)(unsigned char)(signed char)(int)q << (8 * u);
            }
            *reinterpret_cast<unsigned long long*>(out + (long long)i * plane_stride + k8) = pack;
        }
    }
}

// ---- CRT variant: symmetric residues of the scaled rows, one int8 plane per modulus ----
// X as above; planes[i][r][k] = (rint(x 2^(bits-e)) mod p_i) in [-p_i/2, p_i/2),
// scale[r] = 2^(e - bits) so that C = (integer dot product) * scale_row * scale_col.
__global__ void __launch_bounds__(256) ozaki_crt_split_kernel(const double* __restrict__ X,
                                                              long long ld, int K,
                                                              int8_t* __restrict__ planes,
                                                              long long plane_stride,
                                                              double* __restrict__ scale) {
    const int r = blockIdx.x;
    const double* x = X + (long long)r * ld;
    __shared__ double red[256];
    double amax = 0.0;
    for (int k = threadIdx.x; k < K; k += 256) amax = fmax(amax, fabs(x[k]));
    red[threadIdx.x] = amax;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    amax = red[0];
    int e = 0;
    if (amax > 0.0 && isfinite(amax)) e = ilogb(amax) + 1;  // amax * 2^-e in [0.5, 1)
    const int bits = c_crt.bits, nmod = c_crt.nmod;
    if (threadIdx.x == 0) scale[r] = ldexp(1.0, e - bits);
    int8_t* out = planes + (long long)r * K;
    // 8 consecutive k per thread: one 64-byte read, one packed 8-byte store per residue plane
    for (int k8 = threadIdx.x * 8; k8 < K; k8 += 256 * 8) {
        double xi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xi[u] = rint(ldexp(x[k8 + u], bits - e));   // |xi| <= 2^bits
        for (int i = 0; i < nmod; ++i) {
            const int pi = c_crt.p[i];
            const double ii = c_crt.inv_p[i];
            unsigned long long pack = 0ULL;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pack |= (unsigned long long)(unsigned char)(signed char)crt_residue(xi[u], pi, ii)
                        << (8 * u);
            *reinterpret_cast<unsigned long long*>(out + (long long)i * plane_stride + k8) = pack;
        }
    }
}

// C[col*ldc + row] = crt_reconstruct(residues of element (row, col)) * scaleA[row] * scaleB[col]
__global__ void __launch_bounds__(256) ozaki_crt_reconstruct_kernel(
    const unsigned char* __restrict__ R, long long ldr, int M, int N,
    const double* __restrict__ scaleA, const double* __restrict__ scaleB, double* __restrict__ C,
    long long ldc) {
    const int row = blockIdx.x * 256 + threadIdx.x;
    const int col = blockIdx.y;
    if (row >= M) return;
    const double v = crt_reconstruct(R + (long long)col * ldr + row, (long long)N * ldr, c_crt);
    C[(long long)col * ldc + row] = v * (scaleA[row] * scaleB[col]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn ozaki_encode() {
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

bool make_plane_map(CUtensorMap* map, const int8_t* base, long long rows, long long K, int slices,
                    int box_rows) {
    EncodeTiledFn enc = ozaki_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)slices};
    cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)K * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)OBK, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool ozaki_applicable(const Context* ctx, long long m, long long n, long long k) {
    // exact int32 accumulation needs (slices) * K * 64^2 < 2^31; TMA needs 16-byte row strides
    if (ctx->ozaki_crt > 0)   // residues |a|, |b| <= 128: K * 2^14 < 2^31; P covers K <= 2^14
        return (k % 16) == 0 && m >= 512 && n >= 512 && k >= 512 && k <= 16384;
    return ctx->ozaki_slices >= 2 && (k % 16) == 0 && m >= 512 && n >= 512 && k >= 512 &&
           (double)ctx->ozaki_slices * (double)k * 4096.0 < 2147483648.0;
}

// uploads the table for `nmod` moduli into constant memory (stream-ordered; only when it changes)
static void ozaki_crt_table(Context* ctx, int nmod) {
    static int loaded = 0;
    if (loaded == nmod) return;
    TNR_CUDA(cudaMemcpyToSymbolAsync(c_crt, &CRT_TABLES[nmod - CRT_MIN_MOD], sizeof(CrtTable), 0,
                                     cudaMemcpyHostToDevice, ctx->stream));
    loaded = nmod;
}

OzakiOperand ozaki_split(Context* ctx, const double* X, long long ld, long long rows, long long K) {
    OzakiOperand o;
    o.rows = rows;
    o.K = K;
    if (ctx->ozaki_crt > 0) {
        o.slices = ctx->ozaki_crt;
        o.crt = true;
        ozaki_crt_table(ctx, o.slices);
        TNR_CUDA(cudaMallocAsync((void**)&o.planes, (size_t)o.slices * rows * K, ctx->stream));
        o.scale = dalloc(ctx, rows);
        ozaki_crt_split_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(X, ld, (int)K, o.planes,
                                                                       rows * K, o.scale);
        TNR_CUDA(cudaGetLastError());
        ctx->ctr.launches++;
        return o;
    }
    o.slices = ctx->ozaki_slices;
    TNR_CUDA(cudaMallocAsync((void**)&o.planes, (size_t)o.slices * rows * K, ctx->stream));
    o.scale = dalloc(ctx, rows);
    ozaki_split_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(X, ld, (int)K, o.slices, o.planes,
                                                               rows * K, o.scale);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
    return o;
}

void ozaki_free(Context* ctx, OzakiOperand& o) {
    if (o.planes) cudaFreeAsync(o.planes, ctx->stream);
    if (o.scale) dfree(ctx, o.scale);
    o.planes = nullptr;
    o.scale = nullptr;
}

// C(m x n, ldc) = A^T B from the digit planes of A^T (m rows) and B^T (n rows)
void ozaki_multiply(Context* ctx, const OzakiOperand& A, const OzakiOperand& B, double* C,
                    long long ldc) {
    TNR_CHECK(A.K == B.K && A.slices == B.slices && A.crt == B.crt, "ozaki_multiply: operand mismatch");
    CUtensorMap mapA, mapB;
    TNR_CHECK(make_plane_map(&mapA, A.planes, A.rows, A.K, A.slices, OBM) &&
                  make_plane_map(&mapB, B.planes, B.rows, B.K, B.slices, OBN),
              "ozaki_multiply: tensor map encoding failed");
    static bool configured = false;
    if (!configured) {
        TNR_CUDA(cudaFuncSetAttribute(ozaki_tile_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OSMEM));
        TNR_CUDA(cudaFuncSetAttribute(ozaki_tile_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OSMEM));
        configured = true;
    }
    dim3 grid((unsigned)(((A.rows + OBM - 1) / OBM) * ((B.rows + OBN - 1) / OBN)));
    OzakiParams p;
    p.C = C; p.ldc = ldc;
    p.scaleA = A.scale; p.scaleB = B.scale;
    p.M = (int)A.rows; p.N = (int)B.rows; p.K = (int)A.K;
    p.S = A.slices;
    p.R = nullptr; p.ldr = 0;
    if (A.crt) {
        // residues of the accumulators: S bytes per element, reconstructed by a second kernel
        ozaki_crt_table(ctx, A.slices);
        p.ldr = A.rows;
        TNR_CUDA(cudaMallocAsync((void**)&p.R, (size_t)A.slices * A.rows * B.rows, ctx->stream));
        ozaki_tile_kernel<true><<<grid, OTHREADS, OSMEM, ctx->stream>>>(mapA, mapB, p);
        TNR_CUDA(cudaGetLastError());
        dim3 rg((unsigned)((A.rows + 255) / 256), (unsigned)B.rows);
        ozaki_crt_reconstruct_kernel<<<rg, 256, 0, ctx->stream>>>(p.R, p.ldr, p.M, p.N, A.scale,
                                                                 B.scale, C, ldc);
        TNR_CUDA(cudaGetLastError());
        TNR_CUDA(cudaFreeAsync(p.R, ctx->stream));
        ctx->ctr.launches++;
    } else {
        ozaki_tile_kernel<false><<<grid, OTHREADS, OSMEM, ctx->stream>>>(mapA, mapB, p);
        TNR_CUDA(cudaGetLastError());
    }
    ctx->ctr.launches++;
    ctx->ctr.ozaki_launches++;
    ctx->ctr.gemm_flops += 2.0 * A.rows * B.rows * (double)A.K;
    ctx->ctr.ozaki_gemms++;
}

}  // namespace tnr
