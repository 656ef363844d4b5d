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

// merged index groups of a strided copy in destination order (the first step of both planners)
struct BulkGroup { long long n, s, d; };
inline std::vector<BulkGroup> merged_groups(int rank, const long long* dims,
                                            const long long* sstride, const long long* dstride) {
    std::vector<BulkGroup> v, m;
    for (int i = 0; i < rank; ++i)
        if (dims[i] != 1) v.push_back({dims[i], sstride[i], dstride[i]});
    std::stable_sort(v.begin(), v.end(),
                     [](const BulkGroup& a, const BulkGroup& b) { return a.d < b.d; });
    for (auto& x : v) {
        if (!m.empty() && m.back().s * m.back().n == x.s && m.back().d * m.back().n == x.d)
            m.back().n *= x.n;
        else
            m.push_back(x);
    }
    return m;
}

// =====================================================================================
// Bulk-async tiled copy (copy_bulk_kernel, csrc/permute.cu): the read side of a tile is done by
// the TMA engine (cp.async.bulk global -> shared, SASS UBLKCP, completion on an mbarrier), so a
// CTA has its whole tile (<= 72 KB) in flight without holding it in registers; the write side
// is a per-thread phase with 16-byte stores where alignment allows.
//
// Index structure after merging:  v  = group that is fastest on BOTH sides (extent n_v, stride 1
// on both sides; n_v = 1 for a true transposition), i' = (i1, i2) the groups that follow v in the
// SOURCE, j' = (j1, j2) the groups that follow v in the DESTINATION, the rest are outer dims.
// Shared-memory tile: [j'][i2][i1][v], row pitch `pitch` doubles (even, pitch/2 odd: rows start
// 16-byte aligned for the bulk copies and the 16-byte row stride is odd in units of banks).
// =====================================================================================
constexpr int BULK_KMAX = 6;          // elements (of VEC doubles) per lane and output run
constexpr int BULK_TILE_ELEMS = 9216; // doubles per tile (72 KB): 3 CTAs per SM
constexpr int BULK_MAX_TPC = 8;       // tiles per CTA
constexpr int BULK_REPITCH = 5;       // double2 per thread held while a dense tile is re-pitched
constexpr int BULK_DENSE_ELEMS = BULK_REPITCH * 256 * 2;   // doubles per dense tile

struct BulkLaneTab {
    int so[BULK_KMAX];        // shared-memory offset of the slot (doubles), -1 = unused
    int ri[BULK_KMAX];        // row of the pass the slot belongs to
    long long dof[BULK_KMAX]; // destination offset relative to the pass origin
};

struct BulkParams {
    int rank;
    long long dims[MAXR], ss[MAXR], ds[MAXR];
    long long n_v, n_i1, n_i2, n_j1, n_j2;
    long long s_i1, s_i2, s_j1, s_j2;
    long long d_i1, d_i2, d_j1, d_j2;
    int TV, TI1, TI2, TJ1, TJ2;
    long long tiles_v, tiles_i1, tiles_i2, tiles_j1, tiles_j2;
    int pitch;
    int contig1, contig2;   // source row of a tile: (v,i1) contiguous / (v,i1,i2) contiguous
    int vec;                // 2: 16-byte shared loads + global stores in the write phase, else 1
    int R;                  // i1 rows written per warp pass (short destination runs are batched)
    int tpc;                // tiles per CTA (small tiles: several in flight per CTA)
    int chunked;            // 1: rows shorter than 512 B are fetched as 16-byte cp.async chunks
    int tab_smem;           // 1: the lane table is computed by one warp and shared (small tiles)
    int dense;              // 1: short source rows that follow each other contiguously are fetched
                            //    as ONE piece per j2 slice and re-pitched in shared memory
    long long ntiles;       // total number of tiles (grid = ceil(ntiles / tpc))
};

struct BulkPlan {
    bool ok = false;
    BulkParams p{};
    long long blocks = 0;
    size_t smem = 0;
};

inline int bulk_pick_extent(long long n, long long tgt, bool want_even) {
    if (tgt < 1) tgt = 1;
    if (n <= tgt) return (int)n;
    for (long long t = tgt; 2 * t >= tgt && t >= 1; --t)
        if (n % t == 0 && (!want_even || t % 2 == 0)) return (int)t;
    if (want_even && tgt > 1 && tgt % 2) --tgt;
    return (int)tgt;
}

// m: merged groups in destination order (m[0] destination-fastest).  Returns ok = false when
// the copy does not have the structure / alignment the bulk kernel needs.
struct BulkTuning {
    int max_tpc = BULK_MAX_TPC;     // tiles per CTA (1 = one tile per CTA)
    int dense = 1;                  // dense fetch + re-pitch of short contiguous rows
    int chunk_below = 0;            // whole-row pieces below this many bytes use cp.async chunks
    //                                 (0 = never: measured equal to bulk pieces, more instructions)
};

inline BulkPlan plan_bulk_copy(const std::vector<BulkGroup>& m, uintptr_t src_addr,
                               uintptr_t dst_addr, int tile_tgt,
                               const BulkTuning& tune = BulkTuning()) {
    BulkPlan out;
    const size_t none = (size_t)-1;
    if (m.empty() || m.size() > (size_t)MAXR + 4) return out;
    if (src_addr % 16 != 0) return out;
    BulkParams& p = out.p;
    std::vector<bool> used(m.size(), false);
    size_t iv = none;
    if (m[0].s == 1 && m[0].d == 1) { iv = 0; used[0] = true; }
    p.n_v = (iv != none) ? m[0].n : 1;
    // i1: smallest source stride among the rest; j1: first of the rest in destination order
    // (the destination-contiguous group keeps priority over a source continuation i2)
    size_t i1 = none;
    for (size_t i = 0; i < m.size(); ++i)
        if (!used[i] && (i1 == none || m[i].s < m[i1].s)) i1 = i;
    if (i1 == none) return out;                      // a flat copy: not ours
    if (iv == none && m[i1].s != 1) return out;      // no unit-stride run in the source
    used[i1] = true;
    size_t j1 = none;
    for (size_t i = 0; i < m.size(); ++i)
        if (!used[i]) { j1 = i; break; }
    if (j1 != none) used[j1] = true;
    size_t i2 = none;
    for (size_t i = 0; i < m.size(); ++i)
        if (!used[i] && m[i].s == m[i1].s * m[i1].n) i2 = i;
    if (i2 != none) used[i2] = true;
    size_t j2 = none;
    if (j1 != none)
        for (size_t i = 0; i < m.size(); ++i)
            if (!used[i] && m[i].d == m[j1].d * m[j1].n) j2 = i;
    if (j2 != none) used[j2] = true;
    auto G = [&](size_t k, long long& n, long long& s, long long& d) {
        if (k == none) { n = 1; s = 0; d = 0; } else { n = m[k].n; s = m[k].s; d = m[k].d; }
    };
    G(i1, p.n_i1, p.s_i1, p.d_i1); G(i2, p.n_i2, p.s_i2, p.d_i2);
    G(j1, p.n_j1, p.s_j1, p.d_j1); G(j2, p.n_j2, p.s_j2, p.d_j2);
    p.rank = 0;
    long long outer = 1;
    for (size_t i = 0; i < m.size(); ++i)
        if (m[i].n >= (1LL << 31)) return out;       // 32-bit index arithmetic in the kernel
    for (size_t i = 0; i < m.size(); ++i) {
        if (used[i]) continue;
        if (p.rank >= MAXR) return out;
        p.dims[p.rank] = m[i].n; p.ss[p.rank] = m[i].s; p.ds[p.rank] = m[i].d;
        p.rank++;
        outer *= m[i].n;
    }
    // ---- tile extents ----------------------------------------------------------------
    const bool dst_even = dst_addr % 16 == 0 && p.d_i1 % 2 == 0 && p.d_i2 % 2 == 0 &&
                          p.d_j1 % 2 == 0 && p.d_j2 % 2 == 0 && [&] {
                              for (int d = 0; d < p.rank; ++d) if (p.ds[d] % 2) return false;
                              return true; }();
    const int run_max1 = 32 * BULK_KMAX;            // doubles per output run with 8-byte stores
    if (p.n_v == 1) {
        p.TV = 1;
        const int tgt = std::min(tile_tgt, 96);
        auto split = [&](long long n1, long long n2, int& T1, int& T2, bool even) {
            if (n1 >= tgt) { T1 = bulk_pick_extent(n1, tgt, even); T2 = 1; }
            else { T1 = (int)n1; T2 = n2 <= 1 ? 1 : bulk_pick_extent(n2, tgt / n1, false); }
        };
        split(p.n_i1, p.n_i2, p.TI1, p.TI2, true);
        split(p.n_j1, p.n_j2, p.TJ1, p.TJ2, false);
        p.vec = 1;
    } else {
        p.vec = (p.n_v % 2 == 0 && dst_even) ? 2 : 1;
        const int run_max = run_max1 * p.vec;
        p.TV = (p.n_v <= run_max) ? (int)p.n_v : bulk_pick_extent(p.n_v, run_max, true);
        if (p.vec == 2 && p.TV % 2) p.vec = 1;
        const long long tgt_j = std::max<long long>(1, (long long)(run_max1 * p.vec) / p.TV);
        if (p.n_j1 >= tgt_j) { p.TJ1 = bulk_pick_extent(p.n_j1, tgt_j, false); p.TJ2 = 1; }
        else { p.TJ1 = (int)p.n_j1;
               p.TJ2 = p.n_j2 <= 1 ? 1 : bulk_pick_extent(p.n_j2, tgt_j / p.n_j1, false); }
        const long long tgt_i = std::max<long long>(1, BULK_TILE_ELEMS /
                                                    ((long long)p.TJ1 * p.TJ2 * p.TV));
        if (p.n_i1 >= tgt_i) { p.TI1 = bulk_pick_extent(p.n_i1, tgt_i, false); p.TI2 = 1; }
        else { p.TI1 = (int)p.n_i1;
               p.TI2 = p.n_i2 <= 1 ? 1 : bulk_pick_extent(p.n_i2, tgt_i / p.n_i1, false); }
    }
    p.contig1 = (p.s_i1 == p.n_v && p.TV == p.n_v) ? 1 : 0;
    p.contig2 = (p.contig1 && p.TI1 == p.n_i1 && (p.n_i2 == 1 || p.s_i2 == p.s_i1 * p.n_i1)) ? 1 : 0;
    if (p.n_i2 == 1) p.contig2 = p.contig1;   // a single i2 slice: the row is the (v,i1) run
    // ---- 16-byte rule of the bulk copies: every piece starts on an even element and has an
    // even length, for full and ragged tiles alike ----------------------------------------
    auto even = [](long long x) { return x % 2 == 0; };
    if (p.contig1) {
        // piece = TV*ti1 (* ti2): ti1 takes the values TI1 and n_i1 % TI1
        if (!even(p.n_v * p.TI1) || !even(p.n_v * (p.n_i1 % p.TI1))) return out;
    } else {
        if (!even(p.TV) || !even(p.n_v % p.TV) || !even(p.s_i1)) return out;
        if (p.TV * 8 < 64) return out;              // pieces below 64 bytes: not worth a bulk copy
    }
    if (!even(p.s_j1) || !even(p.s_j2)) return out;
    if (p.n_i2 > 1 && !even(p.s_i2)) return out;
    for (int d = 0; d < p.rank; ++d) if (!even(p.ss[d])) return out;
    if ((long long)p.TJ1 * p.TJ2 * p.TV > (long long)run_max1 * p.vec) return out;
    {   // rows per pass: about 96 (VEC = 1) / 192 (VEC = 2) doubles per warp pass
        const long long run = (long long)p.TJ1 * p.TJ2 * p.TV;
        long long R = (32LL * p.vec * 3) / run;
        if (R < 1) R = 1;
        if (R > p.TI1) R = p.TI1;
        if (R * run > (long long)run_max1 * p.vec) R = std::max<long long>(1, ((long long)run_max1 * p.vec) / run);
        p.R = (int)R;
    }
    p.tiles_v = (p.n_v + p.TV - 1) / p.TV;
    p.tiles_i1 = (p.n_i1 + p.TI1 - 1) / p.TI1;
    p.tiles_i2 = (p.n_i2 + p.TI2 - 1) / p.TI2;
    p.tiles_j1 = (p.n_j1 + p.TJ1 - 1) / p.TJ1;
    p.tiles_j2 = (p.n_j2 + p.TJ2 - 1) / p.TJ2;
    long long row = (long long)p.TI1 * p.TI2 * p.TV;
    long long pitch = row + (row & 1);
    if ((pitch / 2) % 2 == 0) pitch += 2;
    if (pitch > (1 << 20)) return out;
    p.pitch = (int)pitch;
    const size_t tile_smem = (size_t)p.TJ1 * p.TJ2 * pitch * sizeof(double);
    if (tile_smem > 100 * 1024) return out;
    p.ntiles = p.tiles_v * p.tiles_i1 * p.tiles_i2 * p.tiles_j1 * p.tiles_j2 * outer;
    if (p.ntiles >= (1LL << 31) || p.ntiles <= 0) return out;
    // small tiles: several per CTA, all requested up front (about 72 KB per CTA either way)
    long long tpc = (long long)(BULK_TILE_ELEMS * sizeof(double)) / (long long)tile_smem;
    if (tpc < 1) tpc = 1;
    if (tpc > BULK_MAX_TPC) tpc = BULK_MAX_TPC;
    if (tpc > tune.max_tpc) tpc = std::max(1, tune.max_tpc);
    // keep at least ~4 CTAs per SM worth of blocks
    while (tpc > 1 && p.ntiles / tpc < 148 * 4) --tpc;
    p.tpc = (int)tpc;
    // several tiles per CTA: the per-lane slot table of the write phase is computed by ONE warp
    // and shared (its integer divisions were 65 % issue utilisation on 18 KB tiles)
    p.tab_smem = (tpc > 1 && tile_smem * (size_t)tpc + sizeof(BulkLaneTab) * 32 <=
                                 BULK_TILE_ELEMS * sizeof(double)) ? 1 : 0;
    // whole-row pieces below 512 bytes go through 16-byte cp.async chunks instead of one bulk
    // request per piece
    p.chunked = (p.contig2 && row * 8 < tune.chunk_below) ? 1 : 0;
    // rows of less than 512 bytes that follow each other contiguously in the source (j1 continues
    // the row: a 2-D transposition with a short source-contiguous leg): one bulk request per j2
    // slice instead of one per 192-byte row, then an in-place re-pitch through registers
    p.dense = (tune.dense && !p.chunked && p.contig2 && row * 8 < 512 && p.s_j1 == row &&
               (long long)p.TJ1 * p.TJ2 * row <= BULK_DENSE_ELEMS) ? 1 : 0;
    out.smem = tile_smem * (size_t)tpc + (p.tab_smem ? sizeof(BulkLaneTab) * 32 : 0);
    out.blocks = (p.ntiles + tpc - 1) / tpc;
    out.ok = true;
    return out;
}

struct BulkGeom {
    const double* sp;
    double* dp;
    int tv, ti1, ti2, tj1, tj2, ci, cj;
};

TNR_HD BulkGeom bulk_geometry(const double* src, double* dst, const BulkParams& p, long long bid64) {
    // the planner guarantees ntiles < 2^31 and every extent < 2^31: 32-bit divisions
    unsigned bid = (unsigned)bid64;
    unsigned q;
    q = bid / (unsigned)p.tiles_v;  const unsigned t_v = bid - q * (unsigned)p.tiles_v;   bid = q;
    q = bid / (unsigned)p.tiles_i1; const unsigned t_i1 = bid - q * (unsigned)p.tiles_i1; bid = q;
    q = bid / (unsigned)p.tiles_i2; const unsigned t_i2 = bid - q * (unsigned)p.tiles_i2; bid = q;
    q = bid / (unsigned)p.tiles_j1; const unsigned t_j1 = bid - q * (unsigned)p.tiles_j1; bid = q;
    q = bid / (unsigned)p.tiles_j2; const unsigned t_j2 = bid - q * (unsigned)p.tiles_j2; bid = q;
    long long soff = 0, doff = 0;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
        if (d < p.rank) {
            q = bid / (unsigned)p.dims[d];
            const long long i = bid - q * (unsigned)p.dims[d];
            bid = q;
            soff += i * p.ss[d];
            doff += i * p.ds[d];
        }
    }
    const long long v0 = (long long)t_v * p.TV, i10 = (long long)t_i1 * p.TI1,
                    i20 = (long long)t_i2 * p.TI2, j10 = (long long)t_j1 * p.TJ1,
                    j20 = (long long)t_j2 * p.TJ2;
    BulkGeom g;
    g.tv = (int)((p.n_v - v0 < p.TV) ? p.n_v - v0 : p.TV);
    g.ti1 = (int)((p.n_i1 - i10 < p.TI1) ? p.n_i1 - i10 : p.TI1);
    g.ti2 = (int)((p.n_i2 - i20 < p.TI2) ? p.n_i2 - i20 : p.TI2);
    g.tj1 = (int)((p.n_j1 - j10 < p.TJ1) ? p.n_j1 - j10 : p.TJ1);
    g.tj2 = (int)((p.n_j2 - j20 < p.TJ2) ? p.n_j2 - j20 : p.TJ2);
    g.ci = g.ti1 * g.ti2;
    g.cj = g.tj1 * g.tj2;
    g.sp = src + soff + v0 + i10 * p.s_i1 + i20 * p.s_i2 + j10 * p.s_j1 + j20 * p.s_j2;
    g.dp = dst + doff + v0 + i10 * p.d_i1 + i20 * p.d_i2 + j10 * p.d_j1 + j20 * p.d_j2;
    return g;
}

TNR_HD long long bulk_tile_bytes(const BulkGeom& g) {
    return (long long)g.cj * g.ci * g.tv * 8;
}

TNR_HD long long bulk_tile_elems(const BulkParams& p) {
    return (long long)p.TJ1 * p.TJ2 * p.pitch;
}

// chunked load phase (rows shorter than 512 B, contiguous in the source): every thread fetches
// 16-byte chunks; `issue16(dst, src)` is cp.async (LDGSTS) on the device
template <typename Issue16>
TNR_HD void bulk_load_phase_chunked(const BulkGeom& g, const BulkParams& p, double* tile, int tid,
                                    int nthreads, Issue16 issue16) {
    const unsigned cpr = (unsigned)(g.ci * g.tv) / 2;   // 16-byte chunks per row
    const unsigned total = (unsigned)g.cj * cpr;
    for (unsigned c = (unsigned)tid; c < total; c += (unsigned)nthreads) {
        const unsigned r = c / cpr, o = c - r * cpr;
        const unsigned j2 = (g.cj == g.tj1) ? 0u : r / (unsigned)g.tj1;
        const unsigned j1 = r - j2 * (unsigned)g.tj1;
        issue16(tile + (long long)r * p.pitch + 2 * o,
                g.sp + (long long)j1 * p.s_j1 + (long long)j2 * p.s_j2 + 2 * o);
    }
}

// load phase: the pieces of the tile are dealt round-robin to the threads; `issue(dst, src,
// bytes)` is cp.async.bulk on the device and memcpy in the host check
template <typename Issue>
TNR_HD void bulk_load_phase(const BulkGeom& g, const BulkParams& p, double* tile, int tid,
                            int nthreads, Issue issue) {
    const unsigned np1 = p.contig1 ? 1u : (unsigned)g.ti1;
    const unsigned np2 = p.contig2 ? 1u : (unsigned)g.ti2;
    const unsigned np = np1 * np2;
    const int plen = g.tv * (p.contig1 ? g.ti1 : 1) * (p.contig2 ? g.ti2 : 1);
    const unsigned total = (unsigned)g.cj * np;
    for (unsigned q = (unsigned)tid; q < total; q += (unsigned)nthreads) {
        unsigned r = q, a1 = 0, a2 = 0;
        if (np > 1) {
            r = q / np;
            const unsigned pc = q - r * np;
            a2 = (np1 == 1) ? pc : pc / np1;
            a1 = pc - a2 * np1;
        }
        const unsigned j2 = (g.cj == g.tj1) ? 0u : r / (unsigned)g.tj1;
        const unsigned j1 = r - j2 * (unsigned)g.tj1;
        const double* sp = g.sp + (long long)j1 * p.s_j1 + (long long)j2 * p.s_j2 +
                           (long long)a1 * p.s_i1 + (long long)a2 * p.s_i2;
        double* tp = tile + (long long)r * p.pitch + (int)((a2 * (unsigned)g.ti1 + a1) * (unsigned)g.tv);
        issue(tp, sp, plen * 8);
    }
}

// dense tiles: one piece per j2 slice (tj1 consecutive rows are contiguous in the source),
// landing densely at the start of the tile buffer
template <typename Issue>
TNR_HD void bulk_load_phase_dense(const BulkGeom& g, const BulkParams& p, double* tile, int tid,
                                  int nthreads, Issue issue) {
    const int row = g.ci * g.tv;
    for (int j2 = tid; j2 < g.tj2; j2 += nthreads)
        issue(tile + (long long)j2 * g.tj1 * row, g.sp + (long long)j2 * p.s_j2, g.tj1 * row * 8);
}

// in-place re-pitch of a dense tile (rows of `row` doubles back to back) to the padded pitch:
// every thread first takes its 16-byte units into registers (phase 0), then -- after a CTA
// barrier -- writes them to their padded position (phase 1)
struct BulkD2 { double x, y; };
TNR_HD void bulk_repitch(const BulkGeom& g, const BulkParams& p, double* tile, int tid,
                         int nthreads, BulkD2 (&regs)[BULK_REPITCH], int phase) {
    const unsigned row2 = (unsigned)(g.ci * g.tv) / 2, n2 = (unsigned)g.cj * row2;
    BulkD2* t2 = reinterpret_cast<BulkD2*>(tile);
#pragma unroll
    for (int k = 0; k < BULK_REPITCH; ++k) {
        const unsigned idx = (unsigned)tid + (unsigned)nthreads * k;
        if (idx >= n2) continue;
        if (phase == 0) regs[k] = t2[idx];
        else {
            const unsigned r = idx / row2, o = idx - r * row2;
            t2[r * (unsigned)(p.pitch / 2) + o] = regs[k];
        }
    }
}

// Per-lane slots of the write phase: a warp pass writes R consecutive i1 rows of the tile (same
// i2); the lanes run along (row, destination run (j2, j1, v)) of the pass, VEC doubles per lane
// and slot.  The table depends only on the tile's (tv, tj1, cj), i.e. it is the same for every
// full tile: the kernel computes it once per CTA (warp 0 -> shared memory) instead of once per
// thread and tile -- the integer divisions here were the limiter of small tiles.

TNR_HD void bulk_lane_table(const BulkParams& p, int tv, int tj1, int cj, int lane, int vec,
                            BulkLaneTab& T) {
    const unsigned run = (unsigned)(cj * tv);
    const unsigned R = (unsigned)p.R;
#pragma unroll
    for (int k = 0; k < BULK_KMAX; ++k) {
        const unsigned e = (unsigned)((lane + 32 * k) * vec);
        T.so[k] = -1;
        T.ri[k] = 0;
        T.dof[k] = 0;
        if (e < R * run) {
            const unsigned rr = (R == 1) ? 0u : e / run, x = e - rr * run;
            const unsigned jj = (tv == 1) ? x : x / (unsigned)tv;
            const unsigned v = x - jj * (unsigned)tv;
            const unsigned j2 = (cj == tj1) ? 0u : jj / (unsigned)tj1;
            const unsigned j1 = jj - j2 * (unsigned)tj1;
            T.ri[k] = (int)rr;
            T.so[k] = (int)(jj * (unsigned)p.pitch + rr * (unsigned)tv + v);
            T.dof[k] = (long long)rr * p.d_i1 + (long long)j1 * p.d_j1 + (long long)j2 * p.d_j2 + v;
        }
    }
}

template <int VEC, typename Store>
TNR_HD void bulk_write_phase(const BulkGeom& g, const BulkParams& p, const double* tile, int warp,
                             const BulkLaneTab& T, Store store) {
    const int R = p.R;
    const int blocks1 = (g.ti1 + R - 1) / R;          // passes per i2 slice
    const int passes = blocks1 * g.ti2;
    int i2 = 0, b1 = warp;                            // pass q = i2 * blocks1 + b1
    while (b1 >= blocks1) { b1 -= blocks1; ++i2; }
    for (int q = warp; q < passes; q += 8) {
        const int i10 = b1 * R;
        const int left = g.ti1 - i10;
        double* gp = g.dp + i10 * p.d_i1 + i2 * p.d_i2;
        const double* t = tile + (i2 * g.ti1 + i10) * g.tv;
#pragma unroll
        for (int k = 0; k < BULK_KMAX; ++k)
            if (T.so[k] >= 0 && T.ri[k] < left) store(gp + T.dof[k], t + T.so[k]);
        b1 += 8;
        while (b1 >= blocks1) { b1 -= blocks1; ++i2; }
    }
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
