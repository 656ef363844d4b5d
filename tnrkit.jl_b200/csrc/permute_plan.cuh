// Planner and per-thread phases of the tiled strided copy (csrc/permute.cu), as plain host /
// device code: the planner decides flat / row / tiled execution and fills CopyParams, the two
// phase functions are the body of copy_tiled_mlp_kernel<U>.  Kept in a header so that the SAME
// code is compiled with g++ and run thread by thread on the CPU against numpy.transpose
// (tests/test_permute_host.py) -- the index arithmetic is checked without a GPU.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

#ifndef TNR_HD
#if defined(__CUDACC__)
#define TNR_HD __host__ __device__ __forceinline__
#else
#define TNR_HD inline
#endif
#endif

namespace tnr {

constexpr int MAXR = 8;

struct CopyParams {
    int rank;                // number of "outer" dims (excluding tile dims for the tiled kernel)
    long long dims[MAXR];
    long long ss[MAXR];
    long long ds[MAXR];
    // row kernel: inner dim
    long long ni, si_s, si_d;
    long long total;
    // tiled kernel: composite source-contiguous index i' = (i1, i2) and composite
    // destination-contiguous index j' = (j1, j2)
    long long n_i1, n_i2, n_j1, n_j2;          // full extents
    long long s_i1, d_i1, d_i2, s_i2;          // strides of i1 / i2 (src, dst)
    long long d_j1, s_j1, s_j2, d_j2;          // strides of j1 / j2
    int TI1, TI2, TJ1, TJ2;                    // tile extents
    long long tiles_i1, tiles_i2, tiles_j1, tiles_j2;
    int pitch;
};

enum CopyKind { COPY_NONE = 0, COPY_FLAT, COPY_ROWS, COPY_TILED };

struct CopyPlan {
    CopyKind kind = COPY_NONE;
    CopyParams p{};
    long long total = 0;     // elements
    long long blocks = 0;    // tiled: grid size
    size_t smem = 0;         // tiled: dynamic shared memory
    const char* error = nullptr;
};

// dst[i.dstride] = src[i.sstride]: merge index groups that stay adjacent on both sides, then
// flat copy (one contiguous group), row copy (same fastest group on both sides) or tiled
// transpose over (source-fastest, destination-fastest) composites of at most `tgt` doubles.
inline CopyPlan plan_strided_copy(int rank, const long long* dims, const long long* sstride,
                                  const long long* dstride, int tgt) {
    CopyPlan out;
    struct D { long long n, s, d; };
    std::vector<D> v;
    long long total = 1;
    for (int i = 0; i < rank; ++i) {
        if (dims[i] < 0) { out.error = "strided_copy: negative dim"; return out; }
        total *= dims[i];
        if (dims[i] != 1) v.push_back({dims[i], sstride[i], dstride[i]});
    }
    out.total = total;
    if (total == 0) return out;
    // order by destination stride, then merge groups that are adjacent on both sides
    std::stable_sort(v.begin(), v.end(), [](const D& a, const D& b) { return a.d < b.d; });
    std::vector<D> m;
    for (auto& x : v) {
        if (!m.empty() && m.back().s * m.back().n == x.s && m.back().d * m.back().n == x.d)
            m.back().n *= x.n;
        else
            m.push_back(x);
    }
    if (m.empty()) m.push_back({1, 1, 1});  // single element
    if (m.size() == 1 && m[0].s == 1 && m[0].d == 1) {
        out.kind = COPY_FLAT;
        return out;
    }
    if ((int)m.size() > MAXR + 3) {
        out.error = "strided_copy: too many index groups after merging";
        return out;
    }
    // dst-fastest is m[0]; find src-fastest
    size_t js = 0;
    for (size_t i = 1; i < m.size(); ++i)
        if (m[i].s < m[js].s) js = i;
    CopyParams& p = out.p;
    if (js == 0) {
        if ((int)m.size() - 1 > MAXR) {
            out.error = "strided_copy: too many index groups after merging";
            return out;
        }
        p.ni = m[0].n; p.si_s = m[0].s; p.si_d = m[0].d;
        p.rank = 0;
        for (size_t i = 1; i < m.size(); ++i) {
            p.dims[p.rank] = m[i].n; p.ss[p.rank] = m[i].s; p.ds[p.rank] = m[i].d;
            p.rank++;
        }
        p.total = total;
        out.kind = COPY_ROWS;
        return out;
    }
    // i1 = source-fastest group, j1 = destination-fastest group (m[0]); i2 / j2 = the
    // groups that continue them contiguously in the source / destination, if any
    const size_t none = (size_t)-1;
    size_t i2 = none, j2 = none;
    for (size_t i = 1; i < m.size(); ++i)
        if (i != js && m[i].s == m[js].s * m[js].n) i2 = i;
    for (size_t i = 1; i < m.size(); ++i)
        if (i != js && i != i2 && m[i].d == m[0].d * m[0].n) j2 = i;
    const int TGT = tgt;  // composite run length (doubles)
    auto split = [&](long long n1, long long n2, int& T1, int& T2) {
        if (n1 >= TGT) { T1 = TGT; T2 = 1; }
        else if (n1 > 48 || n2 <= 1) { T1 = (int)std::min<long long>(n1, 48); T2 = 1;
                                       if (n1 <= TGT) T1 = (int)n1; }
        else { T1 = (int)n1; T2 = (int)std::max<long long>(1, std::min<long long>(n2, TGT / n1)); }
    };
    p.n_i1 = m[js].n; p.s_i1 = m[js].s; p.d_i1 = m[js].d;
    p.n_j1 = m[0].n;  p.s_j1 = m[0].s;  p.d_j1 = m[0].d;
    p.n_i2 = (i2 != none) ? m[i2].n : 1; p.s_i2 = (i2 != none) ? m[i2].s : 0;
    p.d_i2 = (i2 != none) ? m[i2].d : 0;
    p.n_j2 = (j2 != none) ? m[j2].n : 1; p.s_j2 = (j2 != none) ? m[j2].s : 0;
    p.d_j2 = (j2 != none) ? m[j2].d : 0;
    split(p.n_i1, p.n_i2, p.TI1, p.TI2);
    split(p.n_j1, p.n_j2, p.TJ1, p.TJ2);
    p.tiles_i1 = (p.n_i1 + p.TI1 - 1) / p.TI1;
    p.tiles_i2 = (p.n_i2 + p.TI2 - 1) / p.TI2;
    p.tiles_j1 = (p.n_j1 + p.TJ1 - 1) / p.TJ1;
    p.tiles_j2 = (p.n_j2 + p.TJ2 - 1) / p.TJ2;
    p.pitch = p.TI1 * p.TI2 + 1;
    if ((p.pitch & 1) == 0) p.pitch += 1;
    p.rank = 0;
    long long outer = 1;
    for (size_t i = 1; i < m.size(); ++i) {
        if (i == js || i == i2 || i == j2) continue;
        if (p.rank >= MAXR) {
            out.error = "strided_copy: too many index groups after merging";
            return out;
        }
        p.dims[p.rank] = m[i].n; p.ss[p.rank] = m[i].s; p.ds[p.rank] = m[i].d;
        p.rank++;
        outer *= m[i].n;
    }
    out.blocks = p.tiles_i1 * p.tiles_i2 * p.tiles_j1 * p.tiles_j2 * outer;
    if (out.blocks >= (1LL << 31)) { out.error = "strided_copy: grid too large"; return out; }
    out.smem = (size_t)p.TJ1 * p.TJ2 * p.pitch * sizeof(double);
    out.kind = COPY_TILED;
    return out;
}

// Row copy on 16-byte elements: possible when the shared fastest run is contiguous and even on
// both sides, every other stride is even and both base addresses are 16-byte aligned.  Rewrites
// the parameters in units of double2 and returns true, or leaves them untouched.
inline bool rows_vectorize(CopyParams& p, uintptr_t src_addr, uintptr_t dst_addr) {
    bool vec = p.si_s == 1 && p.si_d == 1 && p.ni % 2 == 0 && src_addr % 16 == 0 &&
               dst_addr % 16 == 0;
    for (int d = 0; vec && d < p.rank; ++d) vec = p.ss[d] % 2 == 0 && p.ds[d] % 2 == 0;
    if (!vec) return false;
    p.ni /= 2; p.total /= 2;
    for (int d = 0; d < p.rank; ++d) { p.ss[d] /= 2; p.ds[d] /= 2; }
    return true;
}

// ---- per-thread phases of the tiled kernel with U rows of loads in flight -------------------
struct TileGeom {
    const double* sp;
    double* dp;
    int ti1, ti2, tj1, tj2, ci, cj;   // extents of this tile and of its two composites
};

TNR_HD TileGeom tile_geometry(const double* src, double* dst, const CopyParams& p, long long bid) {
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
    TileGeom g;
    g.ti1 = (int)((p.n_i1 - i10 < p.TI1) ? p.n_i1 - i10 : p.TI1);
    g.ti2 = (int)((p.n_i2 - i20 < p.TI2) ? p.n_i2 - i20 : p.TI2);
    g.tj1 = (int)((p.n_j1 - j10 < p.TJ1) ? p.n_j1 - j10 : p.TJ1);
    g.tj2 = (int)((p.n_j2 - j20 < p.TJ2) ? p.n_j2 - j20 : p.TJ2);
    g.ci = g.ti1 * g.ti2;
    g.cj = g.tj1 * g.tj2;
    g.sp = src + soff + i10 * p.s_i1 + i20 * p.s_i2 + j10 * p.s_j1 + j20 * p.s_j2;
    g.dp = dst + doff + i10 * p.d_i1 + i20 * p.d_i2 + j10 * p.d_j1 + j20 * p.d_j2;
    return g;
}

// read phase: one warp per j' row (8 warps), lanes along the source-contiguous composite i'
// (<= 96 doubles = 3 per lane); U rows are loaded before the first store to the tile
template <int U>
TNR_HD void tile_read_phase(const TileGeom& g, const CopyParams& p, double* tile, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int pitch = p.pitch;
    const long long lane_s = lane * p.s_i1, step_s = 32 * p.s_i1;
    int j1 = warp % g.tj1, j2 = warp / g.tj1;
    const int dj1 = 8 % g.tj1, dj2 = 8 / g.tj1;
    for (int r0 = warp; r0 < g.cj; r0 += 8 * U) {
        double v[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + 8 * u;
            if (r < g.cj) {
                const double* gp = g.sp + j1 * p.s_j1 + j2 * p.s_j2 + lane_s;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (lane + 32 * k < g.ci) v[u][k] = gp[k * step_s];
            }
            j1 += dj1; j2 += dj2;
            if (j1 >= g.tj1) { j1 -= g.tj1; ++j2; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + 8 * u;
            if (r < g.cj) {
                double* t = tile + r * pitch + lane;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (lane + 32 * k < g.ci) t[32 * k] = v[u][k];
            }
        }
    }
}

// write phase: one warp per i' row, lanes along the destination-contiguous composite j'
TNR_HD void tile_write_phase(const TileGeom& g, const CopyParams& p, const double* tile, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int pitch = p.pitch;
    const long long lane_d = lane * p.d_j1, step_d = 32 * p.d_j1;
    int i1 = warp % g.ti1, i2 = warp / g.ti1;
    const int di1 = 8 % g.ti1, di2 = 8 / g.ti1;
    for (int r = warp; r < g.ci; r += 8) {
        double* gp = g.dp + i1 * p.d_i1 + i2 * p.d_i2 + lane_d;
        const double* t = tile + lane * pitch + r;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (lane + 32 * k < g.cj) gp[k * step_d] = t[32 * k * pitch];
        i1 += di1; i2 += di2;
        if (i1 >= g.ti1) { i1 -= g.ti1; ++i2; }
    }
}

}  // namespace tnr
