"""numpy emulation of the libtnrcuda PRIMITIVES  --  TEST INFRASTRUCTURE ONLY.

The block-sparse host layer (`tnrkit.jl_b200/symmetric.py`) only sequences C-ABI calls:
strided block copies, one grouped GEMM per contraction, per-sector SVD / eigh, rank selection.
Whether that SEQUENCE is right (sector bookkeeping, arrows, offsets, leg orders, chunking) does not
depend on who executes the calls, so the `-m "not gpu"` suite runs it against this emulation,
which implements each entry point from its contract in `include/tnrcuda.h` with numpy on host
memory.  The `-m gpu` suite runs the same sequences through the real library.

It is never imported by the product: `tnrkit.jl_b200` has no CPU path and raises without a CUDA
device (tests/test_host_logic.py::test_no_cpu_fallback...).  Nothing here is timed or shipped.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

PSEUDOPOW_TOL = np.finfo(float).eps ** 0.75   # btrg.jl:51-60


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return C.cast(p, C.c_void_p).value or 0


def _flat(p, n):
    """n doubles at host address p as a writable numpy view."""
    n = int(n)
    if n == 0:
        return np.zeros(0)
    return np.ctypeslib.as_array((C.c_double * n).from_address(_addr(p)))


def _f(p, dims):
    """Column-major view (first index fastest) of the compact tensor at p."""
    dims = tuple(int(d) for d in dims)
    n = int(np.prod(dims)) if dims else 1
    return _flat(p, n).reshape(dims, order="F")


def _mat(p, rows, cols, ld):
    """rows x cols column-major matrix with leading dimension ld."""
    rows, cols, ld = int(rows), int(cols), int(ld)
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols))
    flat = _flat(p, (cols - 1) * ld + rows)
    return np.lib.stride_tricks.as_strided(flat, (rows, cols), (8, 8 * ld))


def _mode(s, mode, p):
    s = np.asarray(s, dtype=float)
    if mode == 0:
        return s.copy()
    if mode == 1:
        return np.sqrt(s)
    if mode == 2:
        out = s.copy()
        m = ~(s < PSEUDOPOW_TOL)
        out[m] = s[m] ** p
        return out
    raise ValueError("mode")


def _strided(p, dims, strides):
    dims = [int(d) for d in dims]
    strides = [int(s) for s in strides]
    extent = 1 + sum((d - 1) * s for d, s in zip(dims, strides))
    flat = _flat(p, extent)
    return np.lib.stride_tricks.as_strided(flat, dims, [8 * s for s in strides])


class EmulatedContext:
    """Stands in for `_lib.Context` in CPU tests: same `.call(name, *args)` surface."""

    device = 0
    torch_device = "cpu"

    def __init__(self):
        self.calls = {}

    def call(self, name, *args):
        self.calls[name] = self.calls.get(name, 0) + 1
        getattr(self, "_" + name)(*args)

    def check(self, rc, what):
        assert rc == 0, what

    def counters(self):
        return {"grouped_gemm_launches": self.calls.get("tnr_gemm_grouped", 0),
                "launches": sum(self.calls.values())}

    def synchronize(self):
        pass

    # ---- include/tnrcuda.h: primitives -------------------------------------------------
    def _tnr_strided_copy(self, src, dst, rank, dims, sstride, dstride):
        d = [dims[i] for i in range(rank)]
        if any(x == 0 for x in d):
            return
        s = _strided(src, d, [sstride[i] for i in range(rank)])
        t = _strided(dst, d, [dstride[i] for i in range(rank)])
        t[...] = s

    def _tnr_permute(self, src, dst, rank, dims, perm):
        d = [dims[i] for i in range(rank)]
        p = [perm[i] for i in range(rank)]
        a = _f(src, d)
        _f(dst, [d[k] for k in p])[...] = np.transpose(a, p)

    def _tnr_gemm_grouped(self, ta, tb, count, probs, alpha, beta):
        ta, tb = ta.decode().upper(), tb.decode().upper()
        for g in range(count):
            pr = probs[g]
            A = _mat(pr.A, pr.k if ta == "T" else pr.m, pr.m if ta == "T" else pr.k, pr.lda)
            B = _mat(pr.B, pr.n if tb == "T" else pr.k, pr.k if tb == "T" else pr.n, pr.ldb)
            Cm = _mat(pr.C, pr.m, pr.n, pr.ldc)
            prod = (A.T if ta == "T" else A) @ (B.T if tb == "T" else B)
            Cm[...] = alpha * prod + (beta * Cm if beta != 0.0 else 0.0)

    def _tnr_gemm(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cp, ldc):
        ta, tb = ta.decode().upper(), tb.decode().upper()
        Am = _mat(A, k if ta == "T" else m, m if ta == "T" else k, lda)
        Bm = _mat(B, n if tb == "T" else k, k if tb == "T" else n, ldb)
        Cm = _mat(Cp, m, n, ldc)
        prod = (Am.T if ta == "T" else Am) @ (Bm.T if tb == "T" else Bm)
        Cm[...] = alpha * prod + (beta * Cm if beta != 0.0 else 0.0)

    def _tnr_scale(self, x, n, alpha):
        _flat(x, n)[...] *= alpha

    def _tnr_axis_scale(self, A, m1, n, m2, s, mode, p):
        a = _f(A, (m1, n, m2))
        a *= _mode(_flat(s, n), mode, p)[None, :, None]

    def _tnr_vec_map(self, s, out, n, mode, p):
        _flat(out, n)[...] = _mode(_flat(s, n), mode, p)

    def _tnr_topk_select(self, vals, n, k, rank_host, eps_out):
        v = np.abs(_flat(vals, n))
        order = np.argsort(-v, kind="stable")
        ranks = np.empty(n, dtype=np.int64)
        ranks[order] = np.arange(n)
        for j in range(n):
            rank_host[j] = int(ranks[j])
        eps = float(np.sqrt(np.sum(v[ranks >= k] ** 2)))
        if eps_out is not None:
            C.cast(eps_out, C.POINTER(C.c_double))[0] = eps

    def _tnr_strided_sum(self, src, rank, dims, stride, weights, sum_out):
        d = [dims[i] for i in range(rank)]
        a = _strided(src, d, [stride[i] for i in range(rank)]).astype(float)
        if weights:
            for ax in range(rank):
                w = weights[ax]
                if w:
                    sh = [1] * rank
                    sh[ax] = d[ax]
                    a = a * _flat(w, d[ax]).reshape(sh)
        C.cast(sum_out, C.POINTER(C.c_double))[0] = float(a.sum())

    def _tnr_contract(self, A, ra, da, la, B, rb, db, lb, Cp, lc):
        la, lb, lc = la.decode(), lb.decode(), lc.decode()
        a = _f(A, [da[i] for i in range(ra)])
        b = _f(B, [db[i] for i in range(rb)])
        size = dict(zip(la, a.shape))
        size.update(zip(lb, b.shape))
        out = np.einsum(f"{la},{lb}->{lc}", a, b, optimize=True)
        _f(Cp, [size[c] for c in lc])[...] = out

    def _tnr_svd_trunc(self, T, rank, dims, ncod, chi, U, S, Vt, k_out, eps_out):
        d = [dims[i] for i in range(rank)]
        m, n = int(np.prod(d[:ncod])), int(np.prod(d[ncod:]))
        M = _f(T, (m, n))
        u, s, vh = np.linalg.svd(M, full_matrices=False)
        k = min(chi, m, n)
        _f(U, (m, k))[...] = u[:, :k]
        _flat(S, k)[...] = s[:k]
        _f(Vt, (k, n))[...] = vh[:k]
        C.cast(k_out, C.POINTER(C.c_int64))[0] = k
        C.cast(eps_out, C.POINTER(C.c_double))[0] = float(np.linalg.norm(s[k:]))

    def _tnr_orth_r(self, T, rank, dims, ncod, R):
        d = [dims[i] for i in range(rank)]
        m, n = int(np.prod(d[:ncod])), int(np.prod(d[ncod:]))
        assert m >= n, "orth_r: expects a tall matrix"
        _f(R, (n, n))[...] = np.linalg.qr(_f(T, (m, n)), mode="r")

    def _tnr_fill_random(self, x, n, seed):
        """csrc/elementwise.cu: fill_random_kernel (splitmix64 of seed + golden * (i + 1))."""
        n = int(n)
        with np.errstate(over="ignore"):
            z = np.uint64(int(seed) & (2 ** 64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * \
                np.arange(1, n + 1, dtype=np.uint64)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        _flat(x, n)[...] = (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0

    def _tnr_orthonormalize(self, A, m, n, refused_out):
        """csrc/pchol.cu: cholqr2 -- Q = A R^-1 with the Cholesky factor R (positive diagonal) of
        A^T A, refused when the factor's diagonal spans more than 1e5 (or is not positive)."""
        M = _f(A, (int(m), int(n)))
        out = C.cast(refused_out, C.POINTER(C.c_int))
        if int(n) > 152:
            out[0] = 1
            return
        try:
            R = np.linalg.cholesky(M.T @ M).T
        except np.linalg.LinAlgError:
            out[0] = 1
            return
        dg = np.diag(R)
        if not np.all(np.isfinite(dg)) or not dg.min() > 1e-5 * dg.max():
            out[0] = 1
            return
        q, r = np.linalg.qr(M)
        M[...] = q * np.where(np.diag(r) < 0, -1.0, 1.0)
        out[0] = 0

    PSD_REL = 8 * np.finfo(float).eps      # csrc/pchol.cu: PCHOL_REL

    def _tnr_psd_factor(self, G, n, L, rank_out):
        """The algorithm of csrc/pchol.cu restated column by column: diagonal pivoting WITHOUT
        row exchanges (L is not triangular; only L L^T = G matters), stop when the largest
        remaining diagonal entry is <= PSD_REL * max_i G[i, i], |L[i, j]| clamped to
        sqrt(d[i]) (Cauchy-Schwarz of a PSD Schur complement) so that rounding noise can
        never be amplified."""
        n = int(n)
        S = _f(G, (n, n))
        S = 0.5 * (S + S.T)
        Lm = np.zeros((n, n))
        d = np.maximum(np.diag(S).copy(), 0.0)
        thresh = self.PSD_REL * (d.max() if n else 0.0)
        active = np.ones(n, dtype=bool)
        r = 0
        for j in range(n):
            cand = np.where(active, d, -1.0)
            p = int(np.argmax(cand))
            if not cand[p] > thresh:
                break
            piv = np.sqrt(d[p])
            col = (S[:, p] - Lm[:, :j] @ Lm[p, :j]) / piv
            lim = np.sqrt(d)
            col = np.clip(col, -lim, lim)
            col[~active] = 0.0
            col[p] = piv
            Lm[:, j] = col
            d = np.maximum(d - col * col, 0.0)
            active[p] = False
            r = j + 1
        _f(L, (n, n))[...] = Lm
        C.cast(rank_out, C.POINTER(C.c_int64))[0] = r

    def _tnr_eigh_trunc(self, MM, n, chi, W, V, k_out, eps_out):
        M = _f(MM, (n, n))
        w, v = np.linalg.eigh(0.5 * (M + M.T))
        order = np.argsort(-np.abs(w), kind="stable")
        k = min(chi, n)
        _flat(W, k)[...] = w[order[:k]]
        _f(V, (n, k))[...] = v[:, order[:k]]
        C.cast(k_out, C.POINTER(C.c_int64))[0] = k
        C.cast(eps_out, C.POINTER(C.c_double))[0] = float(np.linalg.norm(w[order[k:]]))
