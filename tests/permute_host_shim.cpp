// Host build (g++) of csrc/permute_plan.cuh for tests/test_permute_host.py -- test infrastructure.
// Runs the planner of strided_copy and, for the tiled case, the two per-thread phases of
// copy_tiled_mlp_kernel<U> block by block, thread by thread (phase 1 for all 256 threads, then
// phase 2: the __syncthreads() of the kernel).
#include <cstring>
#include <vector>

#include "../tnrkit.jl_b200/csrc/permute_plan.cuh"

using namespace tnr;

// the bulk kernel's two phases: every "thread" issues its pieces (memcpy stands in for
// cp.async.bulk and CHECKS the 16-byte rules of the instruction), then the write phase runs
static int run_bulk(const double* src, double* dst, const BulkPlan& bp, long long* info) {
    const BulkParams& p = bp.p;
    info[0] = (long long)p.TI1 * p.TI2 * p.TV; info[1] = (long long)p.TJ1 * p.TJ2 * p.TV;
    info[2] = bp.blocks; info[3] = (long long)bp.smem; info[4] = p.pitch; info[5] = p.vec;
    std::vector<double> storage(bp.smem / sizeof(double) + 2);
    double* tile_buf = storage.data();
    if ((uintptr_t)tile_buf % 16) ++tile_buf;
    int bad = 0;
    const long long tile_elems = bulk_tile_elems(p);
    if ((size_t)(tile_elems * p.tpc) * sizeof(double) + (p.tab_smem ? sizeof(BulkLaneTab) * 32 : 0) != bp.smem ||
        p.tpc < 1 || p.tpc > BULK_MAX_TPC)
        return -4;
    info[2] = p.ntiles;
    for (long long b = 0; b < bp.blocks; ++b) {
        std::fill(storage.begin(), storage.end(), -12345.678);
        const long long t0 = b * p.tpc;
        const int nt = (int)std::min<long long>(p.tpc, p.ntiles - t0);
        // load phase of every tile of the CTA first, then the write phases (the kernel's order)
        for (int t = 0; t < nt; ++t) {
            BulkGeom g = bulk_geometry(src, dst, p, t0 + t);
            long long bytes = 0;
            auto issue = [&](double* tp, const double* sp, int nbytes) {
                if (nbytes <= 0 || nbytes % 16 || (uintptr_t)tp % 16 || (uintptr_t)sp % 16) ++bad;
                std::memcpy(tp, sp, (size_t)nbytes);
                bytes += nbytes;
            };
            auto issue16 = [&](double* tp, const double* sp) { issue(tp, sp, 16); };
            double* tb = tile_buf + t * tile_elems;
            for (int tid = 0; tid < 256; ++tid) {
                if (p.chunked) bulk_load_phase_chunked(g, p, tb, tid, 256, issue16);
                else if (p.dense) bulk_load_phase_dense(g, p, tb, tid, 256, issue);
                else bulk_load_phase(g, p, tb, tid, 256, issue);
            }
            if (bytes != bulk_tile_bytes(g)) ++bad;      // the mbarrier's expect_tx count
            if (p.dense) {   // re-pitch: all threads load (phase 0), barrier, all threads store
                std::vector<BulkD2> regs(256 * BULK_REPITCH);
                for (int ph = 0; ph < 2; ++ph)
                    for (int tid = 0; tid < 256; ++tid) {
                        BulkD2(&rg)[BULK_REPITCH] =
                            *reinterpret_cast<BulkD2(*)[BULK_REPITCH]>(&regs[tid * BULK_REPITCH]);
                        bulk_repitch(g, p, tb, tid, 256, rg, ph);
                    }
                info[5] = 100 + p.vec;   // tells the test that the dense path ran
            }
        }
        for (int t = 0; t < nt; ++t) {
            BulkGeom g = bulk_geometry(src, dst, p, t0 + t);
            const double* tb = tile_buf + t * tile_elems;
            auto st1 = [&](double* gp, const double* tt) { *gp = *tt; };
            auto st2 = [&](double* gp, const double* tt) {
                if ((uintptr_t)gp % 16 || (uintptr_t)tt % 16) ++bad;
                gp[0] = tt[0]; gp[1] = tt[1];
            };
            for (int tid = 0; tid < 256; ++tid) {
                BulkLaneTab tab;
                bulk_lane_table(p, g.tv, g.tj1, g.cj, tid & 31, p.vec, tab);
                if (p.vec == 2) bulk_write_phase<2>(g, p, tb, tid >> 5, tab, st2);
                else bulk_write_phase<1>(g, p, tb, tid >> 5, tab, st1);
            }
        }
    }
    return bad ? -3 : 4;
}

static int run_plan(const double* src, double* dst, int rank, const long long* dims,
                    const long long* sst, const long long* dst_st, int unroll, int tile,
                    long long* info) {
    if (unroll < 0) {   // bulk path requested: kind 4 when the planner accepts, else fall through
        long long total = 1;
        for (int i = 0; i < rank; ++i) total *= dims[i];
        auto m = merged_groups(rank, dims, sst, dst_st);
        const bool flat = m.empty() || (m.size() == 1 && m[0].s == 1 && m[0].d == 1);
        if (total > 0 && !flat) {
            BulkPlan bp = plan_bulk_copy(m, (uintptr_t)src, (uintptr_t)dst, tile);
            if (bp.ok) return run_bulk(src, dst, bp, info);
        }
        unroll = 1;
    }
    CopyPlan plan = plan_strided_copy(rank, dims, sst, dst_st, tile);
    if (plan.error) return -1;
    const CopyParams& p = plan.p;
    info[0] = (long long)p.TI1 * p.TI2; info[1] = (long long)p.TJ1 * p.TJ2;
    info[2] = plan.blocks; info[3] = (long long)plan.smem; info[4] = p.pitch; info[5] = p.rank;
    if (plan.total == 0) return 0;
    if (plan.kind == COPY_FLAT) {
        std::memcpy(dst, src, sizeof(double) * (size_t)plan.total);
    } else if (plan.kind == COPY_ROWS) {
        CopyParams q = p;
        const bool vec = unroll > 1 && rows_vectorize(q, (uintptr_t)src, (uintptr_t)dst);
        if (vec) info[5] = -1;   // tells the test that the 16-byte path ran
        // the arithmetic of copy_rows_kernel<T>, one "thread" per element of T
        struct D2 { double x, y; };
        if (vec) {
            const D2* s2 = reinterpret_cast<const D2*>(src);
            D2* d2 = reinterpret_cast<D2*>(dst);
            for (long long idx = 0; idx < q.total; ++idx) {
                long long i = idx % q.ni, rest = idx / q.ni;
                long long soff = i * q.si_s, doff = i * q.si_d;
                for (int dd = 0; dd < q.rank; ++dd) {
                    long long k = rest % q.dims[dd];
                    rest /= q.dims[dd];
                    soff += k * q.ss[dd];
                    doff += k * q.ds[dd];
                }
                d2[doff] = s2[soff];
            }
        } else
        for (long long idx = 0; idx < p.total; ++idx) {
            long long i = idx % p.ni, rest = idx / p.ni;
            long long soff = i * p.si_s, doff = i * p.si_d;
            for (int dd = 0; dd < p.rank; ++dd) {
                long long k = rest % p.dims[dd];
                rest /= p.dims[dd];
                soff += k * p.ss[dd];
                doff += k * p.ds[dd];
            }
            dst[doff] = src[soff];
        }
    } else {
        if (p.TI1 * p.TI2 > 96 || p.TJ1 * p.TJ2 > 96) return -2;   // the kernel's precondition
        std::vector<double> tile_buf(plan.smem / sizeof(double));
        for (long long b = 0; b < plan.blocks; ++b) {
            // poison the tile: a read of an element the read phase did not write shows up
            std::fill(tile_buf.begin(), tile_buf.end(), -12345.678);
            TileGeom g = tile_geometry(src, dst, p, b);
            for (int tid = 0; tid < 256; ++tid) {
                if (unroll >= 4) tile_read_phase<4>(g, p, tile_buf.data(), tid);
                else if (unroll == 2) tile_read_phase<2>(g, p, tile_buf.data(), tid);
                else tile_read_phase<1>(g, p, tile_buf.data(), tid);
            }
            for (int tid = 0; tid < 256; ++tid) tile_write_phase(g, p, tile_buf.data(), tid);
        }
    }
    return (int)plan.kind;
}

extern "C" {
// TensorKit-style permute: new leg k = old leg perm[k]; column-major data.
// Returns the plan kind (1 flat, 2 rows, 3 tiled) or a negative number on error;
// info[0..5] = TI1*TI2, TJ1*TJ2, blocks, smem bytes, pitch, merged outer rank.
int permute_host(const double* src, double* dst, int rank, const long long* dims, const int* perm,
                 int unroll, int tile, long long* info) {
    long long sst[16], dims_out[16], dst_st[16], src_st_for_out[16];
    long long s = 1;
    for (int i = 0; i < rank; ++i) { sst[i] = s; s *= dims[i]; }
    long long d = 1;
    for (int k = 0; k < rank; ++k) {
        int q = perm[k];
        dims_out[k] = dims[q];
        dst_st[k] = d;
        d *= dims[q];
        src_st_for_out[k] = sst[q];
    }
    return run_plan(src, dst, rank, dims_out, src_st_for_out, dst_st, unroll, tile, info);
}
// tnr_strided_copy: dst[sum i_k dstride_k] = src[sum i_k sstride_k] (element strides)
int strided_copy_host(const double* src, double* dst, int rank, const long long* dims,
                      const long long* sstride, const long long* dstride, int unroll, int tile,
                      long long* info) {
    return run_plan(src, dst, rank, dims, sstride, dstride, unroll, tile, info);
}
}
