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

namespace tnr {
namespace {

constexpr int MAXR = 8;

struct CopyParams {
    int rank;                // number of "outer" dims (excluding tile dims for the tiled kernel)
    long long dims[MAXR];
    long long ss[MAXR];
    long long ds[MAXR];
    // tile dims (tiled kernel) / inner dim (row kernel)
    long long ni, si_s, si_d;  // src-fastest dim: extent, src stride, dst stride
    long long nj, sj_s, sj_d;  // dst-fastest dim
    long long tiles_i, tiles_j;
    int ti, tj;                // tile extents along i and j
    // slice batch: nb consecutive values of a third index per block
    long long nbdim, sb_s, sb_d, tiles_b;
    int nb;
    long long total;           // total elements (row kernel)
};

constexpr int TMAX = 48;

__global__ void __launch_bounds__(256) copy_tiled_kernel(const double* __restrict__ src,
                                                         double* __restrict__ dst,
                                                         const CopyParams p) {
    extern __shared__ double tile[];  // [nb][TJ][TI + 1]
    long long bid = blockIdx.x;
    long long ti = bid % p.tiles_i;
    bid /= p.tiles_i;
    long long tj = bid % p.tiles_j;
    bid /= p.tiles_j;
    long long tb = bid % p.tiles_b;
    bid /= p.tiles_b;
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
    const int TI = p.ti, TJ = p.tj;
    const long long i0 = ti * TI, j0 = tj * TJ, b0 = tb * p.nb;
    const int ni = (int)min((long long)TI, p.ni - i0), nj = (int)min((long long)TJ, p.nj - j0);
    const int nb = (int)min((long long)p.nb, p.nbdim - b0);
    const double* sp = src + soff + i0 * p.si_s + j0 * p.sj_s + b0 * p.sb_s;
    double* dp = dst + doff + i0 * p.si_d + j0 * p.sj_d + b0 * p.sb_d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = TI + 1, slab = TJ * pitch;
    // read: one warp per (j, b) row, lanes along i (source contiguous).  No div/mod in the
    // loops: (b, j) advance by pointer increments.
    const long long lane_s = lane * p.si_s, step_s = 32 * p.si_s;
    for (int b = 0; b < nb; ++b) {
        const double* gb = sp + b * p.sb_s + lane_s;
        double* tb_ = tile + b * slab;
#pragma unroll 3
        for (int j = warp; j < nj; j += 8) {
            const double* g = gb + j * p.sj_s;
            double* t = tb_ + j * pitch;
            long long off = 0;
            for (int i = lane; i < ni; i += 32, off += step_s) t[i] = g[off];
        }
    }
    __syncthreads();
    // write: one warp per (i, b) row, lanes along j (destination contiguous)
    const long long lane_d = lane * p.sj_d, step_d = 32 * p.sj_d;
    for (int b = 0; b < nb; ++b) {
        double* gb = dp + b * p.sb_d + lane_d;
        const double* tb_ = tile + b * slab + lane * pitch;
#pragma unroll 3
        for (int i = warp; i < ni; i += 8) {
            double* g = gb + i * p.si_d;
            const double* t = tb_ + i;
            long long off = 0;
            int toff = 0;
            for (int j = lane; j < nj; j += 32, off += step_d, toff += 32 * pitch) g[off] = t[toff];
        }
    }
}

// same fastest index on both sides: one thread per element, inner index fastest
__global__ void __launch_bounds__(256) copy_rows_kernel(const double* __restrict__ src,
                                                        double* __restrict__ dst,
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
    struct D { long long n, s, d; };
    std::vector<D> v;
    long long total = 1;
    for (int i = 0; i < rank; ++i) {
        TNR_CHECK(dims[i] >= 0, "strided_copy: negative dim");
        total *= dims[i];
        if (dims[i] != 1) v.push_back({dims[i], sstride[i], dstride[i]});
    }
    if (total == 0) return;
    // order by destination stride, then merge groups that are adjacent on both sides
    std::stable_sort(v.begin(), v.end(), [](const D& a, const D& b) { return a.d < b.d; });
    std::vector<D> m;
    for (auto& x : v) {
        if (!m.empty() && m.back().s * m.back().n == x.s && m.back().d * m.back().n == x.d)
            m.back().n *= x.n;
        else
            m.push_back(x);
    }
    ctx->ctr.permute_bytes += 16.0 * (double)total;
    if (m.empty()) {  // single element
        m.push_back({1, 1, 1});
    }
    if (m.size() == 1 && m[0].s == 1 && m[0].d == 1) {
        int vec = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
        long long work = vec ? (total + 1) / 2 : total;
        copy_flat_kernel<<<(unsigned)((work + 255) / 256), 256, 0, ctx->stream>>>(src, dst, total, vec);
        TNR_CUDA(cudaGetLastError());
        ctx->ctr.launches++;
        return;
    }
    TNR_CHECK((int)m.size() <= MAXR + 3, "strided_copy: too many index groups after merging");
    // dst-fastest is m[0]; find src-fastest
    size_t js = 0;
    for (size_t i = 1; i < m.size(); ++i)
        if (m[i].s < m[js].s) js = i;
    CopyParams p{};
    if (js == 0) {
        p.ni = m[0].n; p.si_s = m[0].s; p.si_d = m[0].d;
        p.rank = 0;
        for (size_t i = 1; i < m.size(); ++i) {
            p.dims[p.rank] = m[i].n; p.ss[p.rank] = m[i].s; p.ds[p.rank] = m[i].d;
            p.rank++;
        }
        p.total = total;
        copy_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, dst, p);
    } else {
        p.ni = m[js].n; p.si_s = m[js].s; p.si_d = m[js].d;
        p.nj = m[0].n; p.sj_s = m[0].s; p.sj_d = m[0].d;
        p.ti = p.ni <= TMAX ? (int)p.ni : 32;
        p.tj = p.nj <= TMAX ? (int)p.nj : 32;
        // batch index: the remaining group with the smallest source stride (keeps the reads
        // of one block inside few DRAM pages); batch size targets <= 40 KB of shared memory
        size_t jb = (size_t)-1;
        for (size_t i = 1; i < m.size(); ++i) {
            if (i == js) continue;
            if (jb == (size_t)-1 || m[i].s < m[jb].s) jb = i;
        }
        const long long slab_bytes = (long long)p.tj * (p.ti + 1) * 8;
        if (jb != (size_t)-1) {
            p.nbdim = m[jb].n; p.sb_s = m[jb].s; p.sb_d = m[jb].d;
            p.nb = (int)std::max<long long>(1, std::min<long long>(p.nbdim, 40960 / slab_bytes));
        } else {
            p.nbdim = 1; p.sb_s = 0; p.sb_d = 0; p.nb = 1;
        }
        p.tiles_b = (p.nbdim + p.nb - 1) / p.nb;
        p.rank = 0;
        long long outer = 1;
        for (size_t i = 1; i < m.size(); ++i) {
            if (i == js || i == jb) continue;
            p.dims[p.rank] = m[i].n; p.ss[p.rank] = m[i].s; p.ds[p.rank] = m[i].d;
            p.rank++;
            outer *= m[i].n;
        }
        p.tiles_i = (p.ni + p.ti - 1) / p.ti;
        p.tiles_j = (p.nj + p.tj - 1) / p.tj;
        long long blocks = p.tiles_i * p.tiles_j * p.tiles_b * outer;
        TNR_CHECK(blocks < (1LL << 31), "strided_copy: grid too large");
        size_t smem = (size_t)p.nb * slab_bytes;
        copy_tiled_kernel<<<(unsigned)blocks, 256, smem, ctx->stream>>>(src, dst, p);
    }
    TNR_CUDA(cudaGetLastError());
    ctx->ctr.launches++;
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
