// EXPERIMENT (not part of libtnrcuda): INT8 GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, accumulator in TMEM, operands staged by TMA) -- the engine an
// Ozaki-scheme FP64 emulation of the chi^3 x chi^3 x chi^3 chunk contraction would need.
// C_s32[M x N] (column major, ldc) = A_s8[M x K] * B_s8[N x K]^T, both operands K-major.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/i8_tcgen05_probe \
//        tools/i8_tcgen05_probe.cu
// run:   tools/i8_tcgen05_probe [M N K]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BM = 128, BN = 256, BKB = 128, STAGES = 4;
constexpr int A_BYTES = BM * BKB, B_BYTES = BN * BKB, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NTHREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + 128;

__device__ int g_error = 0;

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// bounded wait: a protocol error turns into a trap instead of a hang
__device__ __forceinline__ void mbar_wait(unsigned bar, int parity) {
    for (long long it = 0; it < (1LL << 28); ++it) {
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
    atomicExch(&g_error, 1);
    asm volatile("trap;\n");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1,
                                            unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// K-major, 128-byte swizzle: LBO = 1 (16 B), SBO = 1024 B (8 rows x 128 B), version 1
__device__ __forceinline__ uint64_t make_desc(unsigned smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_i8(unsigned tmem_c, uint64_t adesc, uint64_t bdesc,
                                        unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
i8gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              int32_t* __restrict__ C, int M, int N, int K, long long ldc) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = (unsigned long long*)(smem + (size_t)STAGES * STAGE_BYTES);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const unsigned tmem_full = smem_u32(bars + 2 * STAGES);
    unsigned* tmem_ptr = (unsigned*)(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int KB = (K + BKB - 1) / BKB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         smem_u32(tmem_ptr)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                unsigned full = full0 + 8 * s;
                mbar_expect_tx(full, STAGE_BYTES);
                unsigned dst = smem_u32(smem + (size_t)s * STAGE_BYTES);
                tma_load_2d(dst, &mapA, kb * BKB, m0, full);
                tma_load_2d(dst + A_BYTES, &mapB, kb * BKB, n0, full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = S32, A = B = signed int8, K-major both, N = 256, M = 128
            const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN >> 3) << 17) |
                                   ((unsigned)(BM >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                unsigned a_base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                unsigned b_base = a_base + A_BYTES;
#pragma unroll
                for (int k = 0; k < BKB / 32; ++k) {
                    umma_i8(tmem_base, make_desc(a_base + k * 32), make_desc(b_base + k * 32), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty0 + 8 * s);  // frees the stage when these MMAs retire
            }
            umma_commit(tmem_full);
        }
    } else {
        // epilogue warps 2..5: lane quarter = warp % 4
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int row = m0 + q * 32 + lane;
        for (int c0 = 0; c0 < BN; c0 += 16) {
            unsigned v[16];
            unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                  "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                  "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (row < M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    int col = n0 + c0 + j;
                    if (col < N) C[(long long)col * ldc + row] = (int32_t)v[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                     "n"(TMEM_COLS)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e = (x);                                                         \
        if (e != cudaSuccess) {                                                      \
            printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    int M = argc > 3 ? atoi(argv[1]) : 384, N = argc > 3 ? atoi(argv[2]) : 768,
        K = argc > 3 ? atoi(argv[3]) : 640;
    bool verify = (long long)M * N * K <= (1LL << 31);
    printf("i8 tcgen05 GEMM probe: M=%d N=%d K=%d\n", M, N, K);
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    srand(1);
    for (auto& x : hA) x = (int8_t)(rand() % 129 - 64);
    for (auto& x : hB) x = (int8_t)(rand() % 129 - 64);
    int8_t *dA, *dB;
    int32_t* dC;
    CK(cudaMalloc(&dA, hA.size()));
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    CUtensorMap mapA, mapB;
    auto mk = [&](CUtensorMap* mp, void* base, int rows, int boxrows) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)K};
        cuuint32_t box[2] = {(cuuint32_t)BKB, (cuuint32_t)boxrows};
        cuuint32_t es[2] = {1, 1};
        return enc(mp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (mk(&mapA, dA, M, BM) != CUDA_SUCCESS || mk(&mapB, dB, N, BN) != CUDA_SUCCESS) {
        printf("tensor map encode failed (K must be a multiple of 16)\n");
        return 1;
    }
    CK(cudaFuncSetAttribute(i8gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    i8gemm_kernel<<<grid, NTHREADS, SMEM>>>(mapA, mapB, dC, M, N, K, M);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (verify) {
        std::vector<int32_t> hC((size_t)M * N);
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0, checked = 0;
        for (int n = 0; n < N; n += 7)
            for (int m = 0; m < M; m += 5) {
                long long acc = 0;
                for (int k = 0; k < K; ++k) acc += (int)hA[(size_t)m * K + k] * (int)hB[(size_t)n * K + k];
                ++checked;
                if (acc != hC[(size_t)n * M + m]) {
                    if (bad < 5) printf("  mismatch (%d,%d): got %d want %lld\n", m, n, hC[(size_t)n * M + m], acc);
                    ++bad;
                }
            }
        printf("verify: %lld / %lld sampled entries wrong\n", bad, checked);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        i8gemm_kernel<<<grid, NTHREADS, SMEM>>>(mapA, mapB, dC, M, N, K, M);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("time %.3f ms -> %.1f TOPS\n", best, 2.0 * M * N * (double)K / (best * 1e-3) / 1e12);
    return 0;
}
