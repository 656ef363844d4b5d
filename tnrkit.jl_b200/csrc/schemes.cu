// Scheme-level step! / finalize! bodies on device tensors (dense `Trivial` sector).
// Each function cites the reference body it replaces.  Everything runs on
// ctx->stream; the only host synchronisations are the Jacobi convergence flags
// and the final norm read-back of finalize!.
#include "schemes.hpp"

namespace tnr {

namespace {

// U[..., k] *= f(S[k]) : columns of the (rows x k) matrix
void scale_cols(DT& U, const DT& S, int mode, double p) {
    long long k = U.d.back();
    diag_scale(U.ctx, U.p, U.size() / k, k, U.size() / k, S.p, false, mode, p);
}
// V[k, ...] *= f(S[k]) : rows of the (k x cols) matrix
void scale_rows(DT& V, const DT& S, int mode, double p) {
    long long k = V.d.front();
    diag_scale(V.ctx, V.p, k, V.size() / k, k, S.p, true, mode, p);
}

double read_scalar(Context* ctx, const double* dptr) {
    double h = 0.0;
    TNR_CUDA(cudaMemcpyAsync(&h, dptr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TNR_CUDA(cudaStreamSynchronize(ctx->stream));
    return h;
}

}  // namespace

// ---------------------------------------------------------------------------
// finalize!  (src/utility/finalize.jl:4-14, 56-66)
// ---------------------------------------------------------------------------
double finalize_2d(Context* ctx, DT& T) {
    TNR_CHECK(T.rank() == 4 && T.d[0] == T.d[3] && T.d[1] == T.d[2], "finalize: bad 2D tensor");
    // n = |T[1 2; 2 1]|
    long long s0 = 1, s1 = T.d[0], s2 = s1 * T.d[1], s3 = s2 * T.d[2];
    long long dims[2] = {T.d[0], T.d[1]}, st[2] = {s0 + s3, s1 + s2};
    DT n(ctx, {1});
    strided_sum(ctx, T.p, 2, dims, st, n.p, true);
    scale_inv_dev(ctx, T.p, T.size(), n.p);
    return read_scalar(ctx, n.p);
}

double finalize_btrg(Context* ctx, DT& T, const DT& S1, const DT& S2) {
    TNR_CHECK(T.rank() == 4 && T.d[0] == T.d[3] && T.d[1] == T.d[2], "finalize: bad 2D tensor");
    // n = |T[1 2; 4 3] S1[4; 2] S2[3; 1]| with diagonal bond tensors: i4 = i2, i3 = i1
    long long s0 = 1, s1 = T.d[0], s2 = s1 * T.d[1], s3 = s2 * T.d[2];
    long long dims[2] = {T.d[0], T.d[1]}, st[2] = {s0 + s3, s1 + s2};
    const double* w[2] = {S2.p, S1.p};
    DT n(ctx, {1});
    strided_sum_w(ctx, T.p, 2, dims, st, w, n.p, true);
    scale_inv_dev(ctx, T.p, T.size(), n.p);
    return read_scalar(ctx, n.p);
}

double finalize_3d(Context* ctx, DT& T) {
    TNR_CHECK(T.rank() == 6 && T.d[0] == T.d[1] && T.d[2] == T.d[4] && T.d[3] == T.d[5],
              "finalize: bad 3D tensor");
    // n = |T[1 1; 2 3 2 3]|
    long long s[6];
    s[0] = 1;
    for (int i = 1; i < 6; ++i) s[i] = s[i - 1] * T.d[i - 1];
    long long dims[3] = {T.d[0], T.d[2], T.d[3]}, st[3] = {s[0] + s[1], s[2] + s[4], s[3] + s[5]};
    DT n(ctx, {1});
    strided_sum(ctx, T.p, 3, dims, st, n.p, true);
    scale_inv_dev(ctx, T.p, T.size(), n.p);
    return read_scalar(ctx, n.p);
}

// ---------------------------------------------------------------------------
// TRG  (src/schemes/trg.jl:38-44, SVD12 in src/utility/projectors.jl:213-219)
// ---------------------------------------------------------------------------
DT trg_step(Context* ctx, const DT& T, int chi) {
    Trunc f = svd_trunc(T, 2, chi);
    scale_cols(f.U, f.S, 1, 0);   // A = U*sqrt(s)   [a b k]
    scale_rows(f.Vt, f.S, 1, 0);  // B = sqrt(s)*V   [k c d]
    DT Tp = permute(T, {1, 3, 0, 2});  // transpose(T, ((2,4),(1,3)))
    Trunc g = svd_trunc(Tp, 2, chi);
    Tp.release();
    scale_cols(g.U, g.S, 1, 0);   // C
    scale_rows(g.Vt, g.S, 1, 0);  // D
    // T[-1 -2;-3 -4] := D[-2;1 2] * B[-1;4 1] * C[4 3;-3] * A[3 2;-4]
    DT X = contract(g.Vt, "bpq", f.Vt, "asp", "bqas");
    DT Y = contract(g.U, "src", f.U, "rqd", "scqd");
    return contract(X, "bqas", Y, "scqd", "abcd");
}

// ---------------------------------------------------------------------------
// BTRG  (src/schemes/btrg.jl:62-97); S1/S2 are kept as their diagonals
// ---------------------------------------------------------------------------
void btrg_step(Context* ctx, DT& T, DT& S1, DT& S2, double kexp, int chi) {
    double pa = (1.0 - kexp) / 2.0;
    Trunc f = svd_trunc(T, 2, chi);
    DT S1n(ctx, f.S.d);
    vec_map(ctx, f.S.p, S1n.p, f.S.size(), 2, kexp);  // S1' = pseudopow(S, k)
    scale_cols(f.U, f.S, 2, pa);                      // A = U * S_a      [6 5; -3] -> "psc"
    scale_rows(f.Vt, f.S, 2, pa);                     // B = S_a * V      [-2; 1 3] -> "bqr"
    DT Tp = permute(T, {2, 0, 3, 1});                 // permute(T, ((3,1),(4,2)))
    Trunc g = svd_trunc(Tp, 2, chi);
    Tp.release();
    DT S2n(ctx, g.S.d);
    vec_map(ctx, g.S.p, S2n.p, g.S.size(), 2, kexp);
    scale_cols(g.U, g.S, 2, pa);   // C [8 2; -4] -> "srd"
    scale_rows(g.Vt, g.S, 2, pa);  // D [-1; 4 7] -> "apq"
    // T := D[-1;4 7] S1[1;7] B[-2;1 3] S2[3;2] C[8 2;-4] S1[8;5] A[6 5;-3] S2[4;6]
    // diagonal S1,S2:  B'[b,q,r] = B[b,q,r] s1[q] s2[r];  A'[p,s,c] = A[p,s,c] s2[p] s1[s]
    {
        DT& B = f.Vt;  // [b q r]
        axis_scale(ctx, B.p, B.d[0], B.d[1], B.d[2], S1.p, 0, 0);
        axis_scale(ctx, B.p, B.d[0] * B.d[1], B.d[2], 1, S2.p, 0, 0);
        DT& A = f.U;  // [p s c]
        axis_scale(ctx, A.p, 1, A.d[0], A.d[1] * A.d[2], S2.p, 0, 0);
        axis_scale(ctx, A.p, A.d[0], A.d[1], A.d[2], S1.p, 0, 0);
    }
    DT X = contract(g.Vt, "apq", f.Vt, "bqr", "apbr");
    DT Y = contract(g.U, "srd", f.U, "psc", "rdpc");
    T = contract(X, "apbr", Y, "rdpc", "abcd");
    S1 = std::move(S1n);
    S2 = std::move(S2n);
}

// ---------------------------------------------------------------------------
// HOTRG  (src/schemes/hotrg.jl)
// ---------------------------------------------------------------------------
namespace {

// picks (U, eps) with the smaller truncation error: `if eps > eps' then U', eps'`
DT pick_projector(Context* ctx, Trunc& a, Trunc& b) {
    DT U(ctx, a.U.d);
    select_copy(ctx, U.p, a.U.p, b.U.p, U.size(), a.eps.p, b.eps.p, nullptr);
    return U;
}

DT hotrg_xproj(Context* ctx, const DT& A1, const DT& A2, int chi) {
    // hotrg.jl:102-106
    Trunc l, r;
    {
        DT X = contract(A2, "aeij", A2, "cfij", "aecf");
        DT Y = contract(A1, "bkel", A1, "dkfl", "bedf");
        DT MM = contract(X, "aecf", Y, "bedf", "abcd");
        l = eigh_trunc(std::move(MM), 2, chi);
    }
    {   // hotrg.jl:110-114
        DT X = contract(A2, "jeia", A2, "jfic", "eafc");
        DT Y = contract(A1, "lkeb", A1, "lkfd", "ebfd");
        DT MM = contract(X, "eafc", Y, "ebfd", "abcd");
        r = eigh_trunc(std::move(MM), 2, chi);
    }
    return pick_projector(ctx, l, r);
}

DT hotrg_yproj(Context* ctx, const DT& A1, const DT& A2, int chi) {
    Trunc l, r;
    {   // hotrg.jl:137-141
        DT X = contract(A1, "iaje", A1, "icjf", "aecf");
        DT Y = contract(A2, "ebkl", A2, "fdkl", "ebfd");
        DT MM = contract(X, "aecf", Y, "ebfd", "abcd");
        l = eigh_trunc(std::move(MM), 2, chi);
    }
    {   // hotrg.jl:145-149
        DT X = contract(A1, "ijae", A1, "ijcf", "aecf");
        DT Y = contract(A2, "ekbl", A2, "fkdl", "ebfd");
        DT MM = contract(X, "aecf", Y, "ebfd", "abcd");
        r = eigh_trunc(std::move(MM), 2, chi);
    }
    return pick_projector(ctx, l, r);
}

}  // namespace

DT hotrg_step(Context* ctx, const DT& T0, int chi) {
    DT Ux = hotrg_xproj(ctx, T0, T0, chi);
    // hotrg.jl:57-58: T := conj(Ux[1 2;-1]) Ux[3 4;-4] A2[1 5;-3 3] A1[2 -2;5 4]
    DT W = contract(Ux, "ija", T0, "imck", "jamck");
    W = contract(W, "jamck", T0, "jbml", "ackbl");
    DT T = contract(W, "ackbl", Ux, "kld", "abcd");
    W.release();
    Ux.release();
    DT Uy = hotrg_yproj(ctx, T, T, chi);
    // hotrg.jl:79-80: T := A1[-1 1;3 5] A2[5 2;4 -4] conj(Uy[1 2;-2]) Uy[3 4;-3]
    W = contract(T, "aikm", Uy, "ijb", "akmjb");
    W = contract(W, "akmjb", T, "mjld", "akbld");
    return contract(W, "akbld", Uy, "klc", "abcd");
}

// ---------------------------------------------------------------------------
// ATRG  (src/schemes/atrg.jl:37-82)
// ---------------------------------------------------------------------------
namespace {
DT atrg_half(Context* ctx, const DT& T, int chi) {
    DT Tp = permute(T, {0, 2, 1, 3});  // ((1,3),(2,4))
    Trunc f = svd_trunc(Tp, 2, chi);   // A = f.U [i1 i3 k], B = f.Vt [k i2 i4]
    Tp.release();
    DT C = clone(f.U), Bs = clone(f.Vt);
    scale_rows(Bs, f.S, 0, 0);  // B = S*B
    scale_cols(C, f.S, 0, 0);   // C = C*S
    // M[-1 -2;-3 -4] := B[-3;1 -4] * C[-1 1;-2] ; then permute(M, ((1,3),(2,4))) = [a c b d]
    DT M = contract(C, "aib", Bs, "cid", "acbd");
    C.release();
    Bs.release();
    Trunc g = svd_trunc(M, 2, chi);  // X = g.U [m1 m3 k], Y = g.Vt [k m2 m4]
    M.release();
    scale_cols(g.U, g.S, 1, 0);
    scale_rows(g.Vt, g.S, 1, 0);
    // Q[-1 -2;-3 -4] := A[3 -3;2] * D[1;-2 4] * X[4 2;-4] * Y[-1 1;3]
    DT AX = contract(f.U, "kcj", g.U, "ljd", "kcld");
    DT YD = contract(g.Vt, "aik", f.Vt, "ibl", "akbl");
    DT Q = contract(YD, "akbl", AX, "kcld", "abcd");
    AX.release();
    YD.release();
    Trunc h = svd_trunc(Q, 2, chi);  // H = h.U, G = h.Vt
    scale_cols(h.U, h.S, 1, 0);
    scale_rows(h.Vt, h.S, 1, 0);
    // T[-1 -2;-3 -4] := G[-1;-3 1] * H[1 -2;-4]
    return contract(h.Vt, "aci", h.U, "ibd", "abcd");
}
}  // namespace

DT atrg_step(Context* ctx, const DT& T0, int chi) {
    DT T = atrg_half(ctx, T0, chi);
    T = permute(T, {1, 3, 0, 2});  // ((2,4),(1,3))
    T = atrg_half(ctx, T, chi);
    return permute(T, {2, 0, 3, 1});  // ((3,1),(4,2))
}

// ---------------------------------------------------------------------------
// HOTRG_3D  (src/schemes/hotrg3d.jl)
// ---------------------------------------------------------------------------
namespace {

// MM[x1 x2; x1' x2'] for the open leg `o` (hotrg3d.jl:57-61 with o = 5, :78-82 with o = 3;
// the y-projector permutation ((1,2),(4,3,6,5)) maps them to o = 4 and o = 2).
DT hotrg3d_mm(Context* ctx, const DT& A1, const DT& A2, int o) {
    std::string base = "zwYXyx";  // legs 0..5
    // m2: contract all legs of A2 except 0 and o ; m1: all legs of A1 except 1 and o
    std::string la2 = base, lb2 = base, la1 = base, lb1 = base;
    lb2[0] = 'v'; lb2[o] = 'p';
    lb1[1] = 'v'; lb1[o] = 'p';
    std::string oc(1, base[o]);
    DT m2 = contract(A2, la2, A2, lb2, std::string("z") + oc + "vp");  // [z x2 z' x2']
    DT m1 = contract(A1, la1, A1, lb1, std::string("w") + oc + "vp");  // [z x1 z' x1']
    return contract(m1, "zavc", m2, "zbvd", "abcd");
}

DT hotrg3d_proj(Context* ctx, const DT& A1, const DT& A2, int o_left, int o_right, int chi) {
    Trunc l = eigh_trunc(hotrg3d_mm(ctx, A1, A2, o_left), 2, chi);
    Trunc r = eigh_trunc(hotrg3d_mm(ctx, A1, A2, o_right), 2, chi);
    return pick_projector(ctx, l, r);
}

}  // namespace

Trunc hotrg3d_proj_half(Context* ctx, const DT& T, int which, int chi) {
    static const int open_leg[4] = {5, 3, 4, 2};
    TNR_CHECK(which >= 0 && which < 4, "hotrg3d_proj_half: which in 0..3");
    return eigh_trunc(hotrg3d_mm(ctx, T, T, open_leg[which]), 2, chi);
}

DT hotrg3d_pick(Context* ctx, const DT& Ul, const double* eps_l, const DT& Ur,
                const double* eps_r) {
    TNR_CHECK(Ul.d == Ur.d, "hotrg3d_pick: projector halves differ in shape");
    DT U(ctx, Ul.d);
    select_copy(ctx, U.p, Ul.p, Ur.p, U.size(), eps_l, eps_r, nullptr);
    return U;
}

Dims hotrg3d_substep_dims(const Dims& d, int chi) {
    long long nx = std::min<long long>(chi, d[5] * d[5]);
    long long ny = std::min<long long>(chi, d[4] * d[4]);
    return {d[0], d[1], ny, nx, ny, nx};
}

// One z-compression (hotrg3d.jl:124-129) writing T_out[..., f] for f in [f0, f1).
void hotrg3d_substep(Context* ctx, const DT& T, int chi, DT& Tout, long long f0, long long f1,
                     double* const* peers, int npeers) {
    TNR_CHECK(T.rank() == 6, "hotrg3d: rank-6 tensor expected");
    TNR_CHECK(T.d[0] == T.d[1] && T.d[2] == T.d[4] && T.d[3] == T.d[5], "hotrg3d: leg dims");
    Dims od = hotrg3d_substep_dims(T.d, chi);
    TNR_CHECK(Tout.d == od, "hotrg3d: output dims");
    const long long Dz = T.d[0], Dy = T.d[2], Dx = T.d[3];
    const long long ny = od[2], nx = od[3];
    PhaseScope ph_proj(ctx, "hotrg3d.projectors");
    DT Ux = hotrg3d_proj(ctx, T, T, 5, 3, chi);  // [x1 x2 nx]
    DT Uy = hotrg3d_proj(ctx, T, T, 4, 2, chi);  // [y1 y2 ny]
    ph_proj.stop();
    hotrg3d_contract(ctx, T, Ux, Uy, Tout, f0, f1, peers, npeers);
}

// hotrg3d.jl:116-120 with the projectors given, for the slices f0 <= f < f1 of the new x-bond
void hotrg3d_contract(Context* ctx, const DT& T, const DT& Ux, const DT& Uy, DT& Tout,
                      long long f0, long long f1, double* const* peers, int npeers) {
    const Dims& od = Tout.d;
    const long long Dz = T.d[0], Dy = T.d[2], Dx = T.d[3];
    const long long ny = od[2], nx = od[3];
    TNR_CHECK(Ux.d == Dims({Dx, Dx, nx}) && Uy.d == Dims({Dy, Dy, ny}), "hotrg3d: projector dims");

    //  T[a b c d e f] = Ux[x1 x2 f] Ux[x1' x2' d] Uy[y1 y2 e] Uy[y1' y2' c]
    //                   A1[a z y1' x1' y1 x1] A2[z b y2' x2' y2 x2]
    // chunked over the open bonds (f, d):  R_fd = Qk_f^T Pk_d  is a (Dz Dy Dy)^2 x (Dz Dx Dx) GEMM
    //  Qk_f[(z x1' x2), (a y1' y1)] = sum_x1  A1 * Ux[:, :, f]
    //  Pk_d[(z x1' x2), (y2 b y2')] = sum_x2' A2 * Ux[:, :, d]
    const long long kdim = Dz * Dx * Dx, mdim = Dz * Dy * Dy;
    // opt-in: the chunk GEMM on the INT8 tensor cores (Ozaki scheme, gemm_ozaki.cu); operands
    // are split into digit planes once per d / once per f and reused by all chunks
    const bool oz = ozaki_applicable(ctx, mdim, mdim, kdim);
    // The absorbed operands Pk_d (kdim x mdim doubles each, nx of them: O(chi^7) memory) are
    // built for a WINDOW of d at a time, capped by a byte budget and by the free device memory;
    // Qk_f is rebuilt per window (one chi^7 contraction against the window's nx_w chi^9 GEMMs).
    const double pk_bytes = (double)kdim * (double)mdim * 8.0 * (oz ? 2.0 : 1.0);
    size_t free_b = 0, total_b = 0;
    TNR_CUDA(cudaMemGetInfo(&free_b, &total_b));
    {   // blocks cached by the stream-ordered pool are reusable as well
        cudaMemPool_t pool = nullptr;
        unsigned long long reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess &&
            reserved > used)
            free_b += (size_t)(reserved - used);
        else
            (void)cudaGetLastError();
    }
    // working set besides the window: A2p + Pn + Qn + Qk (4 operand-sized) + R, S, W
    const double reserve = 5.0 * pk_bytes + 3.0 * (double)mdim * (double)mdim * 8.0;
    double budget = std::min((double)ctx->hotrg3d_pk_budget, 0.9 * (double)free_b - reserve);
    long long win = (long long)std::floor(budget / pk_bytes);
    TNR_CHECK(win >= 1,
              "hotrg3d: not enough device memory for one absorbed operand Pk_d plus the working "
              "set (" + std::to_string((pk_bytes + reserve) / 1e9) + " GB needed, " +
                  std::to_string((double)free_b / 1e9) + " GB free)");
    if (win > nx) win = nx;
    const long long so[6] = {1, od[0], od[0] * od[1], od[0] * od[1] * od[2],
                             od[0] * od[1] * od[2] * od[3], od[0] * od[1] * od[2] * od[3] * od[4]};
    DT A2p = permute(T, {3, 0, 1, 2, 4, 5});  // [X | z b Y y x]
    for (long long d0 = 0; d0 < nx; d0 += win) {
    const long long d1 = std::min(nx, d0 + win);
    std::vector<DT> Pk;
    std::vector<OzakiOperand> Pk8;
    Pk.reserve(d1 - d0);
    {
        PhaseScope ph(ctx, "hotrg3d.pk_build");
        for (long long d = d0; d < d1; ++d) {
            DT Uxd = DT::view(ctx, Ux.p + d * Dx * Dx, {Dx, Dx});  // [x1' x2']
            DT Pn = contract(A2p, "XzbYyx", Uxd, "pX", "zbYyxp");
            DT Pd = permute(Pn, {0, 5, 4, 3, 1, 2});  // [z p x | y b Y]
            if (oz) Pk8.push_back(ozaki_split(ctx, Pd.p, kdim, mdim, kdim));
            else Pk.push_back(std::move(Pd));
        }
    }
    if (d1 == nx) A2p.release();
    for (long long f = f0; f < f1; ++f) {
        PhaseScope ph_q(ctx, "hotrg3d.q_build");
        DT Uxf = DT::view(ctx, Ux.p + f * Dx * Dx, {Dx, Dx});  // [x1 x2]
        DT Qn = contract(T, "azYXyx", Uxf, "xq", "azYXyq");
        DT Qk = permute(Qn, {1, 3, 5, 0, 2, 4});  // [z X q | a Y y]
        Qn.release();
        OzakiOperand Q8;
        if (oz) {
            Q8 = ozaki_split(ctx, Qk.p, kdim, mdim, kdim);
            Qk.release();
        }
        ph_q.stop();
        for (long long d = d0; d < d1; ++d) {
            // R[(a y1' y1), (y2 b y2')]
            DT R(ctx, {Dz * Dy, Dy * Dy, Dz * Dy});
            if (oz) {
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                if (ctx->time_gemm) {
                    TNR_CUDA(cudaEventCreate(&e0));
                    TNR_CUDA(cudaEventCreate(&e1));
                    TNR_CUDA(cudaEventRecord(e0, ctx->stream));
                }
                ozaki_multiply(ctx, Q8, Pk8[d - d0], R.p, mdim);
                if (ctx->time_gemm) {
                    TNR_CUDA(cudaEventRecord(e1, ctx->stream));
                    ctx->gemm_events.emplace_back(e0, e1);
                    ctx->timed_flops += 2.0 * mdim * mdim * (double)kdim;
                }
            } else
            gemm(ctx, 'T', 'N', (int)mdim, (int)mdim, (int)kdim, 1.0, Qk.p, kdim, Pk[d - d0].p, kdim,
                 0.0, R.p, mdim);
            PhaseScope ph_uy(ctx, "hotrg3d.uy_scatter");
            // S[(a y1'), e, (b y2')] = sum_(y1 y2) R[(a y1'), (y1 y2), (b y2')] Uy[(y1 y2), e]
            DT S(ctx, {Dz, Dy, ny, Dz, Dy});
            GemmBatch bt;
            bt.nb1 = (int)(Dz * Dy);
            bt.sA1 = Dz * Dy * Dy * Dy;
            bt.sC1 = Dz * Dy * ny;
            gemm(ctx, 'N', 'N', (int)(Dz * Dy), (int)ny, (int)(Dy * Dy), 1.0, R.p, Dz * Dy, Uy.p,
                 Dy * Dy, 0.0, S.p, Dz * Dy, bt);
            R.release();
            // W[a e b c] = sum_(y1' y2') S[a y1' e b y2'] Uy[y1' y2' c]
            DT W = contract(S, "aYebZ", Uy, "YZc", "aebc");
            S.release();
            long long wd[4] = {Dz, ny, Dz, ny};
            long long ws[4] = {1, Dz, Dz * ny, Dz * ny * Dz};
            long long wdst[4] = {so[0], so[4], so[1], so[2]};
            const long long off = d * so[3] + f * so[5];
            if (npeers > 1) {
                // publish the slab to every rank's T' buffer (peer-mapped, NVLink stores)
                double* dst[16];
                for (int r = 0; r < npeers; ++r) dst[r] = peers[r] + off;
                strided_copy_multi(ctx, W.p, dst, npeers, 4, wd, ws, wdst);
            } else {
                strided_copy(ctx, W.p, Tout.p + off, 4, wd, ws, wdst);
            }
        }
        if (oz) ozaki_free(ctx, Q8);
    }
    for (auto& o : Pk8) ozaki_free(ctx, o);
    }
}

// full step! (hotrg3d.jl:131-139) on one GPU
DT hotrg3d_step(Context* ctx, const DT& T0, int chi) {
    DT T = clone(T0);
    for (int it = 0; it < 3; ++it) {
        DT Tn(ctx, hotrg3d_substep_dims(T.d, chi));
        hotrg3d_substep(ctx, T, chi, Tn, 0, Tn.d[5], nullptr, 0);
        T = permute(Tn, {5, 3, 1, 2, 0, 4});  // ((6,4),(2,3,1,5))
    }
    return T;
}

// ---------------------------------------------------------------------------
// ATRG_3D  (src/schemes/atrg3d.jl:34-83)
// ---------------------------------------------------------------------------
namespace {

// Proj_a = Rr * V' * S^-1/2  [pair; k],  Proj_b = S^-1/2 * U' * Rl  [k; pair]   (atrg3d.jl:58-66)
void atrg3d_projectors(Context* ctx, const DT& Rl, const DT& Rr, int chi, DT& Pa, DT& Pb) {
    DT t = contract(Rl, "ip", Rr, "pj", "ij");
    Trunc s = svd_trunc(t, 1, chi);
    scale_cols(s.U, s.S, 2, -0.5);   // U * inv_s   [i k]
    scale_rows(s.Vt, s.S, 2, -0.5);  // inv_s * V   [k j]
    Pa = contract(Rr, "pj", s.Vt, "kj", "pk");
    Pb = contract(s.U, "ik", Rl, "ip", "kp");
}

DT atrg3d_substep(Context* ctx, const DT& T, int chi) {
    std::vector<int> perm = {1, 4, 5, 2, 3, 0};  // ((2,5,6),(3,4,1))
    DT Tp = permute(T, perm);
    Trunc f = svd_trunc(Tp, 3, chi);  // U [i2 i5 i6 k], V [k i3 i4 i1]
    Tp.release();
    // A = permute(U,((4,1),(2,3))) etc. are expressed through labels instead of data movement
    DT US = clone(f.U), SV = clone(f.Vt);
    scale_cols(US, f.S, 0, 0);
    scale_rows(SV, f.S, 0, 0);
    // M[-1 -2;-3 -4 -5 -6] := B[1 -2;-3 -4] * C[-1 1;-5 -6];  B = [i1 k i3 i4] from SV[k i3 i4 i1]
    //   C = [k i2 i5 i6] from US[i2 i5 i6 k]:  M[a b c d e f] = sum_i SV[b c d i] US[i e f a]
    //   we need permute(M, perm) = [b e f c d a]
    DT Mp = contract(US, "iefa", SV, "bcdi", "befcda");
    US.release();
    SV.release();
    Trunc g = svd_trunc(Mp, 3, chi);  // U [m2 m5 m6 k], V [k m3 m4 m1]
    Mp.release();
    scale_cols(g.U, g.S, 1, 0);   // X = [k m2 m5 m6] as g.U[m2 m5 m6 k]
    scale_rows(g.Vt, g.S, 1, 0);  // Y = [m1 k m3 m4] as g.Vt[k m3 m4 m1]
    // AX[-1 -2;-3 -4 -5 -6] := A[1 -2;-3 -5] * X[-1 1;-4 -6]: A[i b c e] = f.U[b c e i],
    //   X[a i d f] = g.U[i d f a]
    DT AX = contract(g.U, "idfa", f.U, "bcei", "abcdef");
    // YD := Y[1 -2;-3 -5] * D[-1 1;-4 -6]: Y[i b c e] = g.Vt[b c e i], D[a i d f] = f.Vt[i d f a]
    DT YD = contract(f.Vt, "idfa", g.Vt, "bcei", "abcdef");
    // R factors (left_orth / right_orth, atrg3d.jl:53-56) up to an orthogonal gauge
    DT R1, R2t, R3, R4t;
    {
        R1 = orth_r(YD, 4);                                 // [r; (5 6)]
        R2t = orth_r(AX, 4);                                // R2 = R2t^T : [(5 6); r]
        DT YDq = permute(YD, {0, 1, 4, 5, 2, 3});
        R3 = orth_r(YDq, 4);                                // [r; (3 4)]
        DT AXq = permute(AX, {0, 1, 4, 5, 2, 3});
        R4t = orth_r(AXq, 4);
    }
    DT R2 = permute(R2t, {1, 0}), R4 = permute(R4t, {1, 0});
    DT P1, P2, P3, P4;
    atrg3d_projectors(ctx, R1, R2, chi, P1, P2);
    atrg3d_projectors(ctx, R3, R4, chi, P3, P4);
    const long long d3 = AX.d[2], d4 = AX.d[3], d5 = AX.d[4], d6 = AX.d[5];
    DT P1t = DT::view(ctx, P1.p, {d5, d6, P1.d[1]});
    DT P2t = DT::view(ctx, P2.p, {P2.d[0], d5, d6});
    DT P3t = DT::view(ctx, P3.p, {d3, d4, P3.d[1]});
    DT P4t = DT::view(ctx, P4.p, {P4.d[0], d3, d4});
    // H[-1 -2;-3 -4] := YD[-1 -2;1 2 3 4] Proj_3[1 2;-3] Proj_1[3 4;-4]
    DT H = contract(contract(YD, "abijkl", P3t, "ijc", "abklc"), "abklc", P1t, "kld", "abcd");
    // G[-1 -2;-3 -4] := AX[-1 -2;1 2 3 4] Proj_4[-3;1 2] Proj_2[-4;3 4]
    DT G = contract(contract(AX, "abijkl", P4t, "cij", "abklc"), "abklc", P2t, "dkl", "abcd");
    // T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] * H[-1 1;-3 -4]
    return contract(H, "aicd", G, "ibef", "abcdef");
}

}  // namespace

DT atrg3d_step(Context* ctx, const DT& T0, int chi) {
    DT T = clone(T0);
    for (int it = 0; it < 3; ++it) {
        T = atrg3d_substep(ctx, T, chi);
        T = permute(T, {3, 5, 1, 4, 0, 2});  // ((4,6),(2,5,1,3))
    }
    return T;
}

}  // namespace tnr
