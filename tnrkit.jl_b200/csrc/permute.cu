// Index permutation / strided copy kernels (HBM-bound).
//
// Replaces TensorKit `permute` / `transpose` (StridedViews copies) used between
// the contractions of every step! body (e.g. /root/reference/src/schemes/hotrg3d.jl:134,
// atrg.jl:40, atrg3d.jl:37).
//
// A permutation is a strided copy dst[i.dstride] = src[i.sstride].  After merging
// index groups that stay adjacent, either the fastest index is the same on both
// sides (row copy) or it differs, in which case a tile over (src-fastest, dst-fastest)
// -- whole legs when the bond dimension is <= 48, 32 x 32 otherwise -- times a batch of
// slices of the next index is transposed through padded shared memory, so that global reads
// run along the source's contiguous leg and global writes along the destination's.
#include <algorithm>

#include "common.cuh"
#include "permute_plan.cuh"

namespace tnr {
namespace {

// Tile = (i1 x i2) x (j1 x j2): reads run along the source-contiguous composite i', writes along
// the destination-contiguous composite j' (e.g. 96-element = 768-byte runs on both sides for
// bond dimension 24), transposed through shared memory with an odd pitch (conflict free).
__global__ void __launch_bounds__(256) copy_tiled_kernel(const double* __restrict__ src,
                                                         double* __restrict__ dst,
                                                         const CopyParams p) {
    extern __shared__ double tile[];  // [TJ1*TJ2][pitch]
    long long bid = blockIdx.x;
    long long t_i1 = bid % p.tiles_i1; bid /= p.tiles_i1;
    long long t_i2 = bid % p.tiles_i2; bid /= p.tiles_i2;
    long long t_j1 = bid % p.tiles_j1; bid /= p.tiles_j1;
    long long t_j2 = bid % p.tiles_j2; bid /= p.tiles_j2;
    long long soff = 0, doff = 0;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
        if (d < p.rank) {
            long long i = bid % p.dims[d];
            bid /= p.dims[d];
            soff += i * p.ss[d];
            doff += i * p.ds[d];
        }
    }
    const long long i10 = t_i1 * p.TI1, i20 = t_i2 * p.TI2, j10 = t_j1 * p.TJ1, j20 = t_j2 * p.TJ2;
    const int ti1 = (int)min((long long)p.TI1, p.n_i1 - i10);
    const int ti2 = (int)min((long long)p.TI2, p.n_i2 - i20);
    const int tj1 = (int)min((long long)p.TJ1, p.n_j1 - j10);
    const int tj2 = (int)min((long long)p.TJ2, p.n_j2 - j20);
    const int ci = ti1 * ti2, cj = tj1 * tj2;   // composite extents of this tile
    const double* sp = src + soff + i10 * p.s_i1 + i20 * p.s_i2 + j10 * p.s_j1 + j20 * p.s_j2;
    double* dp = dst + doff + i10 * p.d_i1 + i20 * p.d_i2 + j10 * p.d_j1 + j20 * p.d_j2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = p.pitch;
    // read phase: one warp per j' row, lanes along the source-contiguous composite i'
    {
        const long long lane_s = lane * p.s_i1, step_s = 32 * p.s_i1;
        int j1 = warp % tj1, j2 = warp / tj1;          // one div per warp, then increments
        const int dj1 = 8 % tj1, dj2 = 8 / tj1;
        for (int r = warp; r < cj; r += 8) {
            const double* g = sp + j1 * p.s_j1 + j2 * p.s_j2 + lane_s;
            double* t = tile + r * pitch;
            long long off = 0;
            for (int i = lane; i < ci; i += 32, off += step_s) t[i] = g[off];
            j1 += dj1; j2 += dj2;
            if (j1 >= tj1) { j1 -= tj1; ++j2; }
        }
    }
    __syncthreads();
    // write phase: one warp per i' row, lanes along the destination-contiguous composite j'
    {
        const long long lane_d = lane * p.d_j1, step_d = 32 * p.d_j1;
        int i1 = warp % ti1, i2 = warp / ti1;
        const int di1 = 8 % ti1, di2 = 8 / ti1;
        for (int r = warp; r < ci; r += 8) {
            double* g = dp + i1 * p.d_i1 + i2 * p.d_i2 + lane_d;
            const double* t = tile + lane * pitch + r;
            long long off = 0;
            int toff = 0;
            for (int j = lane; j < cj; j += 32, off += step_d, toff += 32 * pitch) g[off] = t[toff];
            i1 += di1; i2 += di2;
            if (i1 >= ti1) { i1 -= ti1; ++i2; }
        }
    }
}

// copy_tiled_kernel with U rows of the read phase in flight per warp.  In the kernel above every
// thread waits for its one outstanding 8-byte load before it can store it to shared memory
// (768 threads x 8 B per SM in flight: the 48 % of the HBM roof measured in round 1); here a
// thread issues up to 3 U independent loads (composite run <= 96 doubles = 3 per lane) before the
// first store.  Same tiling, same parameters, same write phase.  Opt-in: tnr_set_option
// "permute_unroll" = 2 | 4 (not yet measured on a B200).
template <int U>
__global__ void __launch_bounds__(256) copy_tiled_mlp_kernel(const double* __restrict__ src,
                                                             double* __restrict__ dst,
                                                             const CopyParams p) {
    extern __shared__ double tile[];  // [TJ1*TJ2][pitch]
    const TileGeom g = tile_geometry(src, dst, p, blockIdx.x);
    tile_read_phase<U>(g, p, tile, threadIdx.x);      // permute_plan.cuh (host-checked)
    __syncthreads();
    tile_write_phase(g, p, tile, threadIdx.x);
}

// ---- bulk-async tiled copy --------------------------------------------------------------
// The whole tile of a CTA is requested from the TMA engine at once (cp.async.bulk, one request
// per contiguous source piece, completion counted in bytes on an mbarrier); nothing is staged
// in registers, so 3 resident CTAs keep ~220 KB per SM in flight / in the write phase.  The
// write phase (permute_plan.cuh: bulk_write_phase) stores runs along the destination.
__device__ __forceinline__ unsigned bsmem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
struct BulkIssue {
    unsigned bar;
    __device__ __forceinline__ void operator()(double* tp, const double* sp, int bytes) const {
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
            ::"r"(bsmem_u32(tp)), "l"(sp), "r"(bytes), "r"(bar)
            : "memory");
    }
};
// explicit st.global: the destination pointer travels through shared memory (tile geometry), so
// the compiler would otherwise fall back to generic-space stores
struct Store1 {
    __device__ __forceinline__ void operator()(double* g, const double* t) const {
        const double v = *t;
        asm volatile("st.global.f64 [%0], %1;\n" ::"l"(g), "d"(v) : "memory");
    }
};
struct Store2 {
    __device__ __forceinline__ void operator()(double* g, const double* t) const {
        const double2 v = *reinterpret_cast<const double2*>(t);
        asm volatile("st.global.v2.f64 [%0], {%1, %2};\n" ::"l"(g), "d"(v.x), "d"(v.y) : "memory");
    }
};

struct Issue16 {
    __device__ __forceinline__ void operator()(double* tp, const double* sp) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                     ::"r"(bsmem_u32(tp)), "l"(sp) : "memory");
    }
};

template <int VEC>
__global__ void __launch_bounds__(256, 3) copy_bulk_kernel(const double* __restrict__ src,
                                                        double* __restrict__ dst,
                                                        const BulkParams p) {
    extern __shared__ __align__(16) double dyn_smem[];
    __shared__ unsigned long long bar_storage[BULK_MAX_TPC];
    __shared__ BulkGeom sgeom[BULK_MAX_TPC];
    // dynamic shared memory: [lane table (32 entries, small tiles only)] [tpc tiles]
    BulkLaneTab* stab = reinterpret_cast<BulkLaneTab*>(dyn_smem);
    double* tile = dyn_smem + (p.tab_smem ? (sizeof(BulkLaneTab) * 32) / sizeof(double) : 0);
    const long long t0 = (long long)blockIdx.x * p.tpc;
    const int nt = (int)min((long long)p.tpc, p.ntiles - t0);
    const long long tile_elems = bulk_tile_elems(p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile origins (64-bit divisions) once per CTA, by the first nt lanes of warp 0
    if (warp == 0) {
        if (lane < nt) sgeom[lane] = bulk_geometry(src, dst, p, t0 + lane);
        if (lane == 0 && !p.chunked) {
            for (int t = 0; t < nt; ++t)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n"
                             ::"r"(bsmem_u32(&bar_storage[t])), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
    } else if (warp == 1 && p.tab_smem) {
        BulkLaneTab tl;
        bulk_lane_table(p, p.TV, p.TJ1, p.TJ1 * p.TJ2, lane, VEC, tl);
        stab[lane] = tl;
    }
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
        const BulkGeom g = sgeom[t];
        if (!p.chunked) {
            const unsigned bar = bsmem_u32(&bar_storage[t]);
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                             ::"r"(bar), "r"((int)bulk_tile_bytes(g)) : "memory");
            if (p.dense) bulk_load_phase_dense(g, p, tile + t * tile_elems, threadIdx.x, 256, BulkIssue{bar});
            else bulk_load_phase(g, p, tile + t * tile_elems, threadIdx.x, 256, BulkIssue{bar});
        } else {
            bulk_load_phase_chunked(g, p, tile + t * tile_elems, threadIdx.x, 256, Issue16{});
        }
    }
    // the write-phase slots of a full tile, computed while the loads are in flight
    BulkLaneTab full_tab;
    if (p.tab_smem) full_tab = stab[lane];
    else bulk_lane_table(p, p.TV, p.TJ1, p.TJ1 * p.TJ2, lane, VEC, full_tab);
    if (p.chunked) {
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
    }
    for (int t = 0; t < nt; ++t) {
        if (!p.chunked) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "BWAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra BDONE;\n"
                "bra BWAIT;\n"
                "BDONE:\n"
                "}\n" ::"r"(bsmem_u32(&bar_storage[t])), "r"(0) : "memory");
        }
        const BulkGeom g = sgeom[t];
        if (p.dense) {
            BulkD2 regs[BULK_REPITCH];
            bulk_repitch(g, p, tile + t * tile_elems, threadIdx.x, 256, regs, 0);
            __syncthreads();
            bulk_repitch(g, p, tile + t * tile_elems, threadIdx.x, 256, regs, 1);
            __syncthreads();
        }
        const bool full = g.tv == p.TV && g.tj1 == p.TJ1 && g.cj == p.TJ1 * p.TJ2;
        if (full) {
            if (VEC == 2) bulk_write_phase<2>(g, p, tile + t * tile_elems, warp, full_tab, Store2{});
            else bulk_write_phase<1>(g, p, tile + t * tile_elems, warp, full_tab, Store1{});
        } else {                                  // ragged tile (rare): its own slots
            BulkLaneTab tab;
            bulk_lane_table(p, g.tv, g.tj1, g.cj, lane, VEC, tab);
            if (VEC == 2) bulk_write_phase<2>(g, p, tile + t * tile_elems, warp, tab, Store2{});
            else bulk_write_phase<1>(g, p, tile + t * tile_elems, warp, tab, Store1{});
        }
    }
}

// same fastest index on both sides: one thread per element, inner index fastest.  T = double2
// (opt-in with "permute_unroll" > 1): the inner run is contiguous and even on both sides and every
// other stride is even, so the copy is the same strided copy on 16-byte elements.
template <typename T>
__global__ void __launch_bounds__(256) copy_rows_kernel(const T* __restrict__ src,
                                                        T* __restrict__ dst,
                                                        const CopyParams p) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= p.total) return;
    long long i = idx % p.ni;
    long long rest = idx / p.ni;
    long long soff = i * p.si_s, doff = i * p.si_d;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
        if (d < p.rank) {
            long long k = rest % p.dims[d];
            rest /= p.dims[d];
            soff += k * p.ss[d];
            doff += k * p.ds[d];
        }
    }
    dst[doff] = src[soff];
}

// Same as copy_rows_kernel but every element is stored to `ndst` destinations with identical
// layout: the local buffer and the peer-mapped buffers of the other GPUs (NVLink stores).  Used
// by the sharded HOTRG_3D step to publish each T' slab to all ranks from the kernel that
// produces it (no separate all-gather).
constexpr int MAXDST = 16;
struct DstTable {
    double* p[MAXDST];
    int n;
};

__global__ void __launch_bounds__(256) copy_rows_multi_kernel(const double* __restrict__ src,
                                                              const DstTable dsts,
                                                              const CopyParams p) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= p.total) return;
    long long i = idx % p.ni;
    long long rest = idx / p.ni;
    long long soff = i * p.si_s, doff = i * p.si_d;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
        if (d < p.rank) {
            long long k = rest % p.dims[d];
            rest /= p.dims[d];
            soff += k * p.ss[d];
            doff += k * p.ds[d];
        }
    }
    const double v = src[soff];
    for (int t = 0; t < dsts.n; ++t) dsts.p[t][doff] = v;
}

// contiguous copy, 16-byte vectorised
__global__ void __launch_bounds__(256) copy_flat_kernel(const double* __restrict__ src,
                                                        double* __restrict__ dst, long long n,
                                                        int vec) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (vec) {
        long long n2 = n / 2;
        if (idx < n2)
            reinterpret_cast<double2*>(dst)[idx] = reinterpret_cast<const double2*>(src)[idx];
        if (idx == 0 && (n & 1)) dst[n - 1] = src[n - 1];
    } else if (idx < n) {
        dst[idx] = src[idx];
    }
}

}  // namespace

void strided_copy(Context* ctx, const double* src, double* dst, int rank, const long long* dims,
                  const long long* sstride, const long long* dstride) {
    TNR_CHECK(rank >= 0 && rank <= 16, "strided_copy: rank out of range");
    CopyPlan plan = plan_strided_copy(rank, dims, sstride, dstride, ctx->permute_tile);
    TNR_CHECK(plan.error == nullptr, plan.error ? plan.error : "");
    const long long total = plan.total;
    if (total == 0) return;
    ctx->ctr.permute_bytes += 16.0 * (double)total;
    CopyParams& p = plan.p;
    if (plan.kind != COPY_FLAT && ctx->permute_bulk) {
        // the TMA-fed kernel whenever the copy has 16-byte aligned source pieces
        BulkTuning tune;
        tune.max_tpc = ctx->permute_tpc;
        tune.chunk_below = ctx->permute_chunk_below;
        tune.dense = ctx->permute_dense ? 1 : 0;
        BulkPlan bp = plan_bulk_copy(merged_groups(rank, dims, sstride, dstride), (uintptr_t)src,
                                     (uintptr_t)dst, ctx->permute_tile, tune);
        if (bp.ok) {
            static bool bulk_configured = false;
            if (!bulk_configured) {
                TNR_CUDA(cudaFuncSetAttribute(copy_bulk_kernel<1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                TNR_CUDA(cudaFuncSetAttribute(copy_bulk_kernel<2>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                bulk_configured = true;
            }
            if (bp.p.vec == 2)
                copy_bulk_kernel<2><<<(unsigned)bp.blocks, 256, bp.smem, ctx->stream>>>(src, dst, bp.p);
            else
                copy_bulk_kernel<1><<<(unsigned)bp.blocks, 256, bp.smem, ctx->stream>>>(src, dst, bp.p);
            TNR_CUDA(cudaGetLastError());
            ctx->ctr.launches++;
            ctx->ctr.permute_bulk_launches++;
            return;
        }
    }
    if (plan.kind == COPY_FLAT) {
        int vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
        long long work = vec ? (total + 1) / 2 : total;
        copy_flat_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(src, dst, total, vec);
    } else if (plan.kind == COPY_ROWS) {
        if (ctx->permute_unroll > 1 && rows_vectorize(p, (uintptr_t)src, (uintptr_t)dst)) {
            copy_rows_kernel<double2><<<(unsigned)((p.total + 255) / 256), 256, 0, ctx->stream>>>(
                reinterpret_cast<const double2*>(src), reinterpret_cast<double2*>(dst), p);
        } else {
            copy_rows_kernel<double><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, dst, p);
        }
    } else {
        const long long blocks = plan.blocks;
        const size_t smem = plan.smem;
        static bool configured = false;
        if (!configured) {
            TNR_CUDA(cudaFuncSetAttribute(copy_tiled_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            TNR_CUDA(cudaFuncSetAttribute(copy_tiled_mlp_kernel<2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            TNR_CUDA(cudaFuncSetAttribute(copy_tiled_mlp_kernel<4>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            configured = true;
        }
        // the unrolled kernel holds a composite run in 3 registers per row: needs runs <= 96
        const bool mlp = ctx->permute_unroll > 1 && p.TI1 * p.TI2 <= 96 && p.TJ1 * p.TJ2 <= 96;
        if (mlp && ctx->permute_unroll >= 4)
            copy_tiled_mlp_kernel<4><<<(unsigned)blocks, 256, smem, ctx->stream>>>(src, dst, p);
        else if (mlp)
            copy_tiled_mlp_kernel<2><<<(unsigned)blocks, 256, smem, ctx->stream>>>(src, dst, p);
        else
            copy_tiled_kernel<<<(unsigned)blocks, 256, smem, ctx->stream>>>(src, dst, p);
    }
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
}

// dst_k[i.dstride] = src[i.sstride] for every destination k (element-wise kernel; used for the
// small chi^4 slabs of the sharded HOTRG_3D step, where the destinations are peer GPUs)
void strided_copy_multi(Context* ctx, const double* src, double* const* dsts, int ndst, int rank,
                        const long long* dims, const long long* sstride,
                        const long long* dstride) {
    TNR_CHECK(ndst >= 1 && ndst <= MAXDST, "strided_copy_multi: 1..16 destinations");
    TNR_CHECK(rank >= 1 && rank <= MAXR + 1, "strided_copy_multi: rank out of range");
    CopyParams p{};
    long long total = 1;
    p.ni = dims[0]; p.si_s = sstride[0]; p.si_d = dstride[0];
    p.rank = 0;
    for (int i = 0; i < rank; ++i) total *= dims[i];
    for (int i = 1; i < rank; ++i) {
        p.dims[p.rank] = dims[i]; p.ss[p.rank] = sstride[i]; p.ds[p.rank] = dstride[i];
        p.rank++;
    }
    if (total == 0) return;
    p.total = total;
    DstTable t{};
    t.n = ndst;
    for (int k = 0; k < ndst; ++k) t.p[k] = dsts[k];
    copy_rows_multi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, t, p);
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
    ctx->ctr.peer_scatter_launches += (ndst > 1);
    ctx->ctr.permute_bytes += 8.0 * (double)total * (1 + ndst);
}

void permute(Context* ctx, const double* src, double* dst, int rank, const long long* dims,
             const int* perm) {
    TNR_CHECK(rank >= 1 && rank <= 16, "permute: rank out of range");
    long long sst[16], dims_out[16], dst_st[16], src_st_for_out[16];
    long long s = 1;
    std::vector<bool> seen(rank, false);
    for (int i = 0; i < rank; ++i) { sst[i] = s; s *= dims[i]; }
    long long d = 1;
    for (int k = 0; k < rank; ++k) {
        int q = perm[k];
        TNR_CHECK(q >= 0 && q < rank && !seen[q], "permute: invalid permutation");
        seen[q] = true;
        dims_out[k] = dims[q];
        dst_st[k] = d;
        d *= dims[q];
        src_st_for_out[k] = sst[q];
    }
    strided_copy(ctx, src, dst, rank, dims_out, src_st_for_out, dst_st);
}

}  // namespace tnr
