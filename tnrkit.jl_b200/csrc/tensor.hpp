// Host-side device tensor + pairwise contraction / truncated factorizations.
// Column major (first index fastest), matching Julia / TensorKit dense blocks.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace tnr {

using Dims = std::vector<long long>;

inline long long prod(const Dims& d, size_t a = 0, size_t b = (size_t)-1) {
    long long p = 1;
    const size_t e = b < d.size() ? b : d.size();
    for (size_t i = a; i < e; ++i) p *= d[i];
    return p;
}

// Device tensor: owns stream-ordered memory unless constructed as a view.
struct DT {
    Context* ctx = nullptr;
    double* p = nullptr;
    Dims d;
    bool owns = false;

    DT() = default;
    DT(Context* c, const Dims& dims) : ctx(c), d(dims), owns(true) {
        p = dalloc(c, (size_t)std::max<long long>(prod(dims), 1));
    }
    static DT view(Context* c, double* ptr, const Dims& dims) {
        DT t;
        t.ctx = c; t.p = ptr; t.d = dims; t.owns = false;
        return t;
    }
    DT(const DT&) = delete;
    DT& operator=(const DT&) = delete;
    DT(DT&& o) noexcept { *this = std::move(o); }
    DT& operator=(DT&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; d = std::move(o.d); owns = o.owns;
            o.p = nullptr; o.owns = false;
        }
        return *this;
    }
    ~DT() { release(); }
    void release() {
        if (owns && p) dfree(ctx, p);
        p = nullptr; owns = false;
    }
    long long size() const { return prod(d); }
    int rank() const { return (int)d.size(); }
    DT reshaped(const Dims& nd) && {  // same storage, new shape
        DT t = std::move(*this);
        t.d = nd;
        return t;
    }
};

// ---- qr.cu: blocked Householder QR (DMMA trailing update) ----
struct QRWork {
    long long m = 0, n = 0;
    DT V;     // explicit reflectors, m x n, unit lower trapezoidal (ld = m)
    DT T;     // compact-WY T blocks, 32 x 32 x ceil(n / 32)
    DT tau;   // n
};
// A (m x n, lda, m >= n) in place: R in the upper triangle; reflectors kept in w
void qr_factor(Context* ctx, double* A, long long m, long long n, long long lda, QRWork& w);
void qr_copy_r(Context* ctx, const double* A, long long lda, long long n, double* R, long long ldr);
// Y (m x k, ldy) <- Q Y ;  Y = Q [X; 0] for X n x k
void qr_apply_q(Context* ctx, const QRWork& w, double* Y, long long ldy, long long k);
void qr_q_times(Context* ctx, const QRWork& w, const double* X, long long ldx, long long k,
                double* Y, long long ldy);

void qr_thin(Context* ctx, const double* A, long long m, long long n, double* Q, double* R);

// ---- pchol.cu: G ~= L L^T for a symmetric PSD matrix (diagonally pivoted Cholesky without row
// exchanges, blocked; columns >= the returned numerical rank are zero).  block <= 0: default.
long long psd_factor(Context* ctx, const double* G, long long n, double* L, int block = 0);

// A (m x n, ld m) <- orthonormal basis of its columns by CholeskyQR2 (Gram GEMM, one-CTA Cholesky +
// inverse, GEMM; twice).  false = refused (not safely positive definite), A untouched.
bool cholqr2(Context* ctx, double* A, long long m, long long n);

DT clone(const DT& a);
DT permute(const DT& a, const std::vector<int>& perm);
// out labels = lc; contracted labels = those in both la and lb and not in lc
DT contract(const DT& A, const std::string& la, const DT& B, const std::string& lb,
            const std::string& lc);

struct Trunc {
    DT U;    // [cod..., k]
    DT S;    // [k]   singular values (svd) / signed eigenvalues (eigh)
    DT Vt;   // [k, dom...]  (svd only)
    DT eps;  // [1] device scalar: 2-norm of the discarded values
};
// svd_trunc(T; trunc = truncrank(chi)) with the first ncod legs as codomain
Trunc svd_trunc(const DT& T, int ncod, int chi);
// eigh_trunc!(project_hermitian!(MM); trunc = truncrank(chi)); MM is [cod..., cod...]
Trunc eigh_trunc(DT MM, int ncod, int chi);
// R-equivalent of left_orth (QR): returns R' (min(m,n) x n) with R'^T R' = A^T A, where A is T
// with the first ncod legs as rows.  R' = S V^T differs from the Householder R by a left
// orthogonal factor, which cancels in every use the reference makes of it (atrg3d.jl:53-66).
DT orth_r(const DT& T, int ncod);

}  // namespace tnr
