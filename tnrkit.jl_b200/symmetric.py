"""Block-sparse tensors with an abelian symmetry: Z_N (TensorKit `Z2Irrep`, `ZNIrrep{N}`) or
U(1) (`U1Irrep`; `N = 0`: integer charges without a modulus, half-integer labels stored doubled).

The host keeps the sector / block structure (which charge tuples are allowed, block shapes,
offsets); the data of every block lives in device memory.  This mirrors how TensorKit stores
an abelian `TensorMap`: for every coupled sector c one dense matrix whose rows / columns are
the fusion-tree (= charge tuple) blocks of the codomain / domain.  The three operations the
TRG / BTRG `step!` bodies need are

  * contraction  = re-block both operands into coupled-sector matrices (strided block copies)
                   and multiply ALL sectors in ONE grouped DMMA launch (`tnr_gemm_grouped`);
  * svd_trunc    = SVD of every coupled-sector matrix + sector-GLOBAL truncrank on the device
                   (`tnr_topk_select` over the concatenated spectra);
  * permute      = per-block index permutation (`tnr_permute`).

Conservation law: a block with charges (q_1..q_r) is present iff sum_i sign_i * q_i = 0 mod N
(= 0 exactly for U(1)), sign = +1 for codomain-like legs and -1 for domain-like legs.  All
fusion / braiding symbols of Z_N and U(1) are 1, so no recoupling coefficients appear.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math

import numpy as np

from . import _lib
from .tensor import DeviceTensor, svd_trunc


class Leg:
    """Graded leg: sectors sorted by charge, `dims[q]` states of charge q, arrow `sign`."""

    def __init__(self, sectors, sign):
        items = sorted((int(q), int(d)) for q, d in dict(sectors).items() if int(d) > 0)
        self.charges = tuple(q for q, _ in items)
        self.dims = {q: d for q, d in items}
        self.sign = int(sign)
        off, acc = {}, 0
        for q, d in items:
            off[q] = acc
            acc += d
        self.offsets = off
        self.total = acc

    def flipped(self):
        return Leg(self.dims, -self.sign)

    def same_space(self, other):
        return self.dims == other.dims

    def __repr__(self):
        return f"Leg({self.dims}, sign={self.sign:+d})"


def _colmajor_strides(dims):
    st, acc = [], 1
    for d in dims:
        st.append(acc)
        acc *= d
    return st


class SymTensor:
    """Z_N (N >= 2) or U(1) (N = 0) block-sparse tensor: `blocks[(q_1..q_r)]` is a dense
    DeviceTensor."""

    def __init__(self, N, legs, blocks, ctx=None):
        self.N = int(N)
        if self.N == 1 or self.N < 0:
            raise ValueError("SymTensor: N must be 0 (U(1)) or >= 2 (Z_N)")
        self.legs = list(legs)
        self.blocks = dict(blocks)
        self._ctx = ctx

    def _fuse(self, c: int) -> int:
        """Coupled charge: reduced mod N for Z_N, as is for U(1)."""
        return c % self.N if self.N else c

    @property
    def ctx(self):
        # resolved lazily so that the sector bookkeeping can be used (and tested) without a GPU
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    # ---- structure ---------------------------------------------------------
    def allowed(self, key):
        return self._fuse(sum(l.sign * q for l, q in zip(self.legs, key))) == 0

    def keys(self):
        for key in itertools.product(*[l.charges for l in reversed(self.legs)]):
            key = tuple(reversed(key))  # first leg fastest
            if self.allowed(key):
                yield key

    def block_dims(self, key):
        return tuple(l.dims[q] for l, q in zip(self.legs, key))

    @property
    def dims(self):
        return tuple(l.total for l in self.legs)

    def nnz(self):
        return sum(math.prod(self.block_dims(k)) for k in self.blocks)

    # ---- host <-> device -----------------------------------------------------
    @classmethod
    def from_dense(cls, arr, N, legs, ctx=None, tol=1e-12):
        """Uploads the symmetry-allowed blocks of a dense array given in the charge basis
        (every leg's index range ordered by sector as in `Leg.offsets`)."""
        arr = np.asarray(arr, dtype=np.float64)
        t = cls(N, legs, {}, ctx)
        assert arr.shape == t.dims, (arr.shape, t.dims)
        mask = np.zeros(arr.shape, dtype=bool)
        for key in t.keys():
            sl = tuple(slice(l.offsets[q], l.offsets[q] + l.dims[q]) for l, q in zip(legs, key))
            t.blocks[key] = DeviceTensor.from_numpy(arr[sl], None, t.ctx)
            mask[sl] = True
        viol = np.abs(arr[~mask]).max() if (~mask).any() else 0.0
        if viol > tol * max(1.0, np.abs(arr).max()):
            raise ValueError(f"tensor is not {'Z%d' % N if N else 'U(1)'} symmetric "
                             f"(forbidden entries up to {viol:.2e})")
        return t

    def to_dense(self):
        out = np.zeros(self.dims)
        for key, blk in self.blocks.items():
            sl = tuple(slice(l.offsets[q], l.offsets[q] + l.dims[q])
                       for l, q in zip(self.legs, key))
            out[sl] = blk.to_numpy()
        return out

    # ---- permute -------------------------------------------------------------
    def permute(self, perm):
        legs = [self.legs[p] for p in perm]
        blocks = {tuple(k[p] for p in perm): b.permute(perm) for k, b in self.blocks.items()}
        return SymTensor(self.N, legs, blocks, self._ctx)

    # ---- coupled-sector matrices ----------------------------------------------
    def _tuples(self, which, negate):
        """{c: [(tuple, offset, size)]} for the legs `which`; c = (+/-) sum sign*q mod N."""
        out = {}
        legs = [self.legs[i] for i in which]
        for key in itertools.product(*[l.charges for l in reversed(legs)]):
            key = tuple(reversed(key))
            c = self._fuse(sum(l.sign * q for l, q in zip(legs, key)))
            if negate:
                c = self._fuse(-c)
            lst = out.setdefault(c, [])
            off = lst[-1][1] + lst[-1][2] if lst else 0
            lst.append((key, off, math.prod(l.dims[q] for l, q in zip(legs, key))))
        return out

    def matricize(self, rows, cols, transposed=False):
        """Coupled-sector matrices M_c (rows_c x cols_c) for the bipartition rows | cols.
        Returns (mats, row_tuples, col_tuples).  transposed=True stores M_c^T (cols_c x rows_c,
        i.e. the COLUMN index contiguous): the K-contiguous operand layout of the TMA GEMM."""
        import torch

        rt, ct = self._tuples(rows, False), self._tuples(cols, True)
        mats = {}
        for c in rt:
            if c not in ct:
                continue
            nr = rt[c][-1][1] + rt[c][-1][2]
            nc = ct[c][-1][1] + ct[c][-1][2]
            buf = torch.zeros(nr * nc, dtype=torch.float64, device=self.ctx.torch_device)
            mats[c] = DeviceTensor(buf, (nc, nr) if transposed else (nr, nc), 1, self.ctx)
        rlook = {c: {k: (o, s) for k, o, s in v} for c, v in rt.items()}
        clook = {c: {k: (o, s) for k, o, s in v} for c, v in ct.items()}
        for key, blk in self.blocks.items():
            rk = tuple(key[i] for i in rows)
            ck = tuple(key[i] for i in cols)
            c = self._fuse(sum(self.legs[i].sign * key[i] for i in rows))
            M = mats[c]
            roff, _ = rlook[c][rk]
            coff, _ = clook[c][ck]
            if transposed:   # the same copy with the roles of the two index groups exchanged
                self._copy_block(blk, key, cols, rows, M, coff, roff, to_matrix=True)
            else:
                self._copy_block(blk, key, rows, cols, M, roff, coff, to_matrix=True)
        return mats, rt, ct

    def _copy_block(self, blk, key, rows, cols, M, roff, coff, to_matrix):
        """Strided copy between a tuple block (compact, leg order of self) and its slot in M."""
        bd = self.block_dims(key)
        sst = _colmajor_strides(bd)
        nr = M.dims[0]
        dst = [0] * len(bd)
        for pos, st in zip(rows, _colmajor_strides([bd[i] for i in rows])):
            dst[pos] = st
        for pos, st in zip(cols, _colmajor_strides([bd[i] for i in cols])):
            dst[pos] = st * nr
        mptr = C.c_void_p(M.buf.data_ptr() + 8 * (roff + coff * nr))
        if to_matrix:
            self.ctx.call("tnr_strided_copy", blk.ptr, mptr, len(bd), _lib.i64(bd), _lib.i64(sst),
                          _lib.i64(dst))
        else:
            self.ctx.call("tnr_strided_copy", mptr, blk.ptr, len(bd), _lib.i64(bd), _lib.i64(dst),
                          _lib.i64(sst))

    # ---- scaling by diagonal (per-sector) weights -------------------------------
    def scale_leg(self, axis, weights, mode=0, p=0.0):
        """T[..., i_axis, ...] *= f(w_{q_axis}[i_axis]) in place; weights: {charge: DeviceTensor}."""
        for key, blk in self.blocks.items():
            bd = self.block_dims(key)
            m1 = math.prod(bd[:axis])
            m2 = math.prod(bd[axis + 1:])
            self.ctx.call("tnr_axis_scale", blk.ptr, m1, bd[axis], m2, weights[key[axis]].ptr,
                          mode, float(p))
        return self

    def scale(self, alpha):
        for blk in self.blocks.values():
            self.ctx.call("tnr_scale", blk.ptr, blk.size, float(alpha))
        return self

    def __repr__(self):
        name = f"Z{self.N}" if self.N else "U1"
        return f"SymTensor({name}, dims={self.dims}, blocks={len(self.blocks)})"


# ------------------------------------------------------------------------------------
def sym_contract(A: SymTensor, la: str, B: SymTensor, lb: str, lc: str) -> SymTensor:
    """One binary `@tensor` contraction of block-sparse tensors: all coupled sectors are
    multiplied in ONE grouped DMMA launch."""
    assert A.N == B.N
    K = [c for c in la if c in lb and c not in lc]
    fa = [c for c in la if c not in K]
    fb = [c for c in lb if c not in K]
    assert sorted(fa + fb) == sorted(lc), "batch / outer labels are not supported"
    for c in K:
        x, y = A.legs[la.index(c)], B.legs[lb.index(c)]
        assert x.sign == -y.sign, f"contracted leg '{c}' must have opposite arrows"
        assert x.same_space(y), f"contracted leg '{c}': sector structure differs"
    rows_a, cols_a = [la.index(c) for c in fa], [la.index(c) for c in K]
    rows_b, cols_b = [lb.index(c) for c in K], [lb.index(c) for c in fb]
    # A is stored transposed (K x M, K contiguous) and B as K x N: both operands K-contiguous, the
    # layout of the TMA + mbarrier GEMM (`tnr_gemm_grouped("T", "N")`: one launch for all sectors)
    Am, art, _ = A.matricize(rows_a, cols_a, transposed=True)
    Bm, _, bct = B.matricize(rows_b, cols_b)
    import torch

    ctx = A.ctx
    Cm, probs = {}, []
    for c in Am:
        if c not in Bm:
            continue
        k, m = Am[c].dims
        k2, n = Bm[c].dims
        assert k == k2
        buf = torch.empty(m * n, dtype=torch.float64, device=ctx.torch_device)
        Cm[c] = DeviceTensor(buf, (m, n), 1, ctx)
        probs.append(_lib.GemmProblem(m, n, k, Am[c].buf.data_ptr(), k, Bm[c].buf.data_ptr(), k,
                                      buf.data_ptr(), m))
    if probs:
        arr = (_lib.GemmProblem * len(probs))(*probs)
        ctx.call("tnr_gemm_grouped", b"T", b"N", len(probs), arr, 1.0, 0.0)
    # split the sector matrices into tuple blocks, directly in the requested leg order
    out_legs_nat = [A.legs[i] for i in rows_a] + [B.legs[i] for i in cols_b]
    nat = fa + fb
    order = [nat.index(c) for c in lc]
    out = SymTensor(A.N, [out_legs_nat[i] for i in order], {}, ctx)
    for c, M in Cm.items():
        nr = M.dims[0]
        for (ka, roff, _), (kb, coff, _) in itertools.product(art[c], bct[c]):
            key_nat = ka + kb
            bd_nat = tuple(l.dims[q] for l, q in zip(out_legs_nat, key_nat))
            src = _colmajor_strides(bd_nat[:len(ka)]) + \
                [s * nr for s in _colmajor_strides(bd_nat[len(ka):])]
            key = tuple(key_nat[i] for i in order)
            bd = tuple(bd_nat[i] for i in order)
            blk = DeviceTensor.empty(bd, None, ctx)
            mptr = C.c_void_p(M.buf.data_ptr() + 8 * (roff + coff * nr))
            ctx.call("tnr_strided_copy", mptr, blk.ptr, len(bd), _lib.i64(bd),
                     _lib.i64([src[i] for i in order]), _lib.i64(_colmajor_strides(bd)))
            out.blocks[key] = blk
    return out


def sym_svd_trunc(T: SymTensor, ncod: int, chi: int):
    """svd_trunc(T; trunc = truncrank(chi)) for a block-sparse tensor: per-sector SVD and a
    sector-global choice of the chi largest singular values (on the device).
    Returns U [cod..., bond(-)], S {c: DeviceTensor}, Vt [bond(+), dom...], eps."""
    import torch

    ctx = T.ctx
    r = len(T.legs)
    rows, cols = list(range(ncod)), list(range(ncod, r))
    mats, rt, ct = T.matricize(rows, cols)
    fac, order = {}, sorted(mats)
    eps2 = 0.0
    for c in order:
        U, S, Vt, e = svd_trunc(mats[c], 1, chi)
        fac[c] = (U, S, Vt)
        eps2 += e * e
    allS = torch.cat([fac[c][1].buf[: fac[c][1].size] for c in order])
    n = allS.numel()
    ranks = (C.c_int32 * n)()
    e_sel = C.c_double()
    ctx.call("tnr_topk_select", C.c_void_p(allS.data_ptr()), n, chi, ranks, C.byref(e_sel))
    eps = math.sqrt(eps2 + e_sel.value ** 2)
    keep, pos = {}, 0
    for c in order:
        ns = fac[c][1].size
        keep[c] = sum(1 for j in range(pos, pos + ns) if ranks[j] < chi)
        pos += ns
    bond = {c: k for c, k in keep.items() if k > 0}
    Ut = SymTensor(T.N, [T.legs[i] for i in rows] + [Leg(bond, -1)], {}, ctx)
    Vt_ = SymTensor(T.N, [Leg(bond, +1)] + [T.legs[i] for i in cols], {}, ctx)
    Sd = {}
    for c, k in bond.items():
        U, S, Vt = fac[c]
        Sd[c] = DeviceTensor(S.buf[:k].clone(), (k,), 1, ctx)
        nr, nc = mats[c].dims
        kc = S.size  # leading dimension of Vt_c
        for key, roff, size in rt[c]:
            bd = tuple(T.legs[i].dims[q] for i, q in zip(rows, key)) + (k,)
            blk = DeviceTensor.empty(bd, None, ctx)
            src = C.c_void_p(U.buf.data_ptr() + 8 * roff)
            ctx.call("tnr_strided_copy", src, blk.ptr, 2, _lib.i64((size, k)), _lib.i64((1, nr)),
                     _lib.i64((1, size)))
            Ut.blocks[key + (c,)] = blk
        for key, coff, size in ct[c]:
            bd = (k,) + tuple(T.legs[i].dims[q] for i, q in zip(cols, key))
            blk = DeviceTensor.empty(bd, None, ctx)
            src = C.c_void_p(Vt.buf.data_ptr() + 8 * coff * kc)
            ctx.call("tnr_strided_copy", src, blk.ptr, 2, _lib.i64((k, size)), _lib.i64((1, kc)),
                     _lib.i64((1, k)))
            Vt_.blocks[(c,) + key] = blk
    return Ut, Sd, Vt_, eps


def vec_map(S, mode, p=0.0):
    """{c: f(S_c)}: sqrt (mode 1) / pseudopow(., p) (mode 2) of per-sector weight vectors."""
    out = {}
    for c, s in S.items():
        o = DeviceTensor.empty(s.dims, 1, s.ctx)
        s.ctx.call("tnr_vec_map", s.ptr, o.ptr, s.size, mode, float(p))
        out[c] = o
    return out


def sym_trace_2d(T: SymTensor, w_leg0=None, w_leg1=None):
    """sum T[1 2; 2 1] (optionally weighted: BTRG's  T[1 2;4 3] S1[4;2] S2[3;1])."""
    total = 0.0
    for key, blk in T.blocks.items():
        q1, q2, q3, q4 = key
        if q3 != q2 or q4 != q1:
            continue
        d = T.block_dims(key)
        assert d[0] == d[3] and d[1] == d[2]
        st = _colmajor_strides(d)
        dims = (d[0], d[1])
        strides = (st[0] + st[3], st[1] + st[2])
        w = (C.c_void_p * 2)(w_leg0[q1].buf.data_ptr() if w_leg0 else None,
                             w_leg1[q2].buf.data_ptr() if w_leg1 else None)
        s = C.c_double()
        T.ctx.call("tnr_strided_sum", blk.ptr, 2, _lib.i64(dims), _lib.i64(strides),
                   w if (w_leg0 or w_leg1) else None, C.byref(s))
        total += s.value
    return total


# ------------------------------------------------------------------------------------
# step! / finalize! bodies on block-sparse tensors (configs[2]: TRG / BTRG on Z2 / ZN)
# ------------------------------------------------------------------------------------
LAST_SPECTRA = {}  # scheme name -> list of {sector: DeviceTensor} kept by the latest step


def trg_step_sym(T: SymTensor, chi: int) -> SymTensor:
    """step!(::TRG) on a Z_N tensor -- src/schemes/trg.jl:38-44 with per-sector SVD12."""
    U, S, V, _ = sym_svd_trunc(T, 2, chi)
    LAST_SPECTRA["trg"] = [S]         # retained per-sector spectra of the step (for inspection)
    rs = vec_map(S, 1)
    A = U.scale_leg(2, rs)            # U * sqrt(s)      [a b k]
    B = V.scale_leg(0, rs)            # sqrt(s) * V      [k c d]
    U2, S2, V2, _ = sym_svd_trunc(T.permute((1, 3, 0, 2)), 2, chi)
    LAST_SPECTRA["trg"].append(S2)
    rs2 = vec_map(S2, 1)
    Cc = U2.scale_leg(2, rs2)
    D = V2.scale_leg(0, rs2)
    # T[-1 -2;-3 -4] := D[-2;1 2] * B[-1;4 1] * C[4 3;-3] * A[3 2;-4]
    X = sym_contract(D, "bpq", B, "asp", "bqas")
    Y = sym_contract(Cc, "src", A, "rqd", "scqd")
    return sym_contract(X, "bqas", Y, "scqd", "abcd")


def btrg_step_sym(T: SymTensor, S1, S2, k: float, chi: int):
    """step!(::BTRG) on a Z_N tensor -- src/schemes/btrg.jl:62-97.  S1, S2: {charge: diag}."""
    pa = (1.0 - k) / 2.0
    U, S, V, _ = sym_svd_trunc(T, 2, chi)
    LAST_SPECTRA["btrg"] = [S]
    Sa, S1n = vec_map(S, 2, pa), vec_map(S, 2, k)
    A = U.scale_leg(2, Sa)            # [p s c]  (btrg labels A[6 5;-3])
    B = V.scale_leg(0, Sa)            # [b q r]
    U2, Sv, V2, _ = sym_svd_trunc(T.permute((2, 0, 3, 1)), 2, chi)
    LAST_SPECTRA["btrg"].append(Sv)
    Sa2, S2n = vec_map(Sv, 2, pa), vec_map(Sv, 2, k)
    Cc = U2.scale_leg(2, Sa2)         # [s r d]
    D = V2.scale_leg(0, Sa2)          # [a p q]
    # T := D[-1;4 7] S1[1;7] B[-2;1 3] S2[3;2] C[8 2;-4] S1[8;5] A[6 5;-3] S2[4;6]
    B.scale_leg(1, S1).scale_leg(2, S2)   # B'[b,q,r] = B s1[q] s2[r]
    A.scale_leg(0, S2).scale_leg(1, S1)   # A'[p,s,c] = A s2[p] s1[s]
    X = sym_contract(D, "apq", B, "bqr", "apbr")
    Y = sym_contract(Cc, "srd", A, "psc", "rdpc")
    return sym_contract(X, "apbr", Y, "rdpc", "abcd"), S1n, S2n


def identity_weights(leg: Leg, ctx):
    return {q: DeviceTensor.from_numpy(np.ones(d), 1, ctx) for q, d in leg.dims.items()}


# ------------------------------------------------------------------------------------
# HOTRG / ATRG on block-sparse tensors
# ------------------------------------------------------------------------------------
def sym_conj(T: SymTensor) -> SymTensor:
    """conj(T) of a real Z_N tensor: same block data, every arrow reversed (dual spaces)."""
    return SymTensor(T.N, [l.flipped() for l in T.legs], dict(T.blocks), T.ctx)


def sym_clone(T: SymTensor) -> SymTensor:
    return SymTensor(T.N, list(T.legs), {k: b.clone() for k, b in T.blocks.items()}, T.ctx)


def sym_eigh_trunc(MM: SymTensor, ncod: int, chi: int):
    """eigh_trunc!(project_hermitian!(MM); trunc = truncrank(chi)) per coupled sector with a
    sector-global choice of the chi eigenvalues of largest magnitude.
    Returns U [cod..., bond(-)], W {c: DeviceTensor}, eps."""
    import torch

    from .tensor import eigh_trunc

    ctx = MM.ctx
    rows, cols = list(range(ncod)), list(range(ncod, len(MM.legs)))
    mats, rt, _ = MM.matricize(rows, cols)
    fac, order, eps2 = {}, sorted(mats), 0.0
    for c in order:
        n = mats[c].dims[0]
        assert mats[c].dims[1] == n, "eigh_trunc: sector blocks must be square"
        W, V, e = eigh_trunc(mats[c], chi)
        fac[c] = (W, V)
        eps2 += e * e
    allW = torch.cat([fac[c][0].buf[: fac[c][0].size] for c in order])
    n = allW.numel()
    ranks = (C.c_int32 * n)()
    e_sel = C.c_double()
    ctx.call("tnr_topk_select", C.c_void_p(allW.data_ptr()), n, chi, ranks, C.byref(e_sel))
    eps = math.sqrt(eps2 + e_sel.value ** 2)
    bond, pos = {}, 0
    for c in order:
        ns = fac[c][0].size
        k = sum(1 for j in range(pos, pos + ns) if ranks[j] < chi)
        pos += ns
        if k > 0:
            bond[c] = k
    U = SymTensor(MM.N, [MM.legs[i] for i in rows] + [Leg(bond, -1)], {}, ctx)
    Wd = {}
    for c, k in bond.items():
        W, V = fac[c]
        Wd[c] = DeviceTensor(W.buf[:k].clone(), (k,), 1, ctx)
        nr = mats[c].dims[0]
        for key, roff, size in rt[c]:
            bd = tuple(MM.legs[i].dims[q] for i, q in zip(rows, key)) + (k,)
            blk = DeviceTensor.empty(bd, None, ctx)
            src = C.c_void_p(V.buf.data_ptr() + 8 * roff)
            ctx.call("tnr_strided_copy", src, blk.ptr, 2, _lib.i64((size, k)), _lib.i64((1, nr)),
                     _lib.i64((1, size)))
            U.blocks[key + (c,)] = blk
    return U, Wd, eps


def _hotrg_proj_sym(mm_left, mm_right, chi):
    """`_, U, eps = eigh_trunc!(MM); ...; if eps > eps' then U', eps'` (hotrg.jl:106-118)."""
    U, _, e = sym_eigh_trunc(mm_left, 2, chi)
    U2, _, e2 = sym_eigh_trunc(mm_right, 2, chi)
    return U2 if e > e2 else U


def hotrg_step_sym(T: SymTensor, chi: int) -> SymTensor:
    """step!(::HOTRG) on a Z_N tensor -- src/schemes/hotrg.jl:155-161."""
    Tc = sym_conj(T)
    # x projector (hotrg.jl:102-114), A1 = A2 = T
    X = sym_contract(T, "aeij", Tc, "cfij", "aecf")
    Y = sym_contract(T, "bkel", Tc, "dkfl", "bedf")
    ML = sym_contract(X, "aecf", Y, "bedf", "abcd")
    X = sym_contract(Tc, "jeia", T, "jfic", "eafc")
    Y = sym_contract(Tc, "lkeb", T, "lkfd", "ebfd")
    MR = sym_contract(X, "eafc", Y, "ebfd", "abcd")
    Ux = _hotrg_proj_sym(ML, MR, chi)
    # T := conj(Ux[1 2;-1]) Ux[3 4;-4] A2[1 5;-3 3] A1[2 -2;5 4]   (hotrg.jl:57-58)
    W = sym_contract(sym_conj(Ux), "ija", T, "imck", "jamck")
    W = sym_contract(W, "jamck", T, "jbml", "ackbl")
    T1 = sym_contract(W, "ackbl", Ux, "kld", "abcd")
    T1c = sym_conj(T1)
    # y projector (hotrg.jl:137-149)
    X = sym_contract(T1, "iaje", T1c, "icjf", "aecf")
    Y = sym_contract(T1, "ebkl", T1c, "fdkl", "ebfd")
    ML = sym_contract(X, "aecf", Y, "ebfd", "abcd")
    X = sym_contract(T1c, "ijae", T1, "ijcf", "aecf")
    Y = sym_contract(T1c, "ekbl", T1, "fkdl", "ebfd")
    MR = sym_contract(X, "aecf", Y, "ebfd", "abcd")
    Uy = _hotrg_proj_sym(ML, MR, chi)
    # T := A1[-1 1;3 5] A2[5 2;4 -4] conj(Uy[1 2;-2]) Uy[3 4;-3]   (hotrg.jl:79-80)
    W = sym_contract(T1, "aikm", sym_conj(Uy), "ijb", "akmjb")
    W = sym_contract(W, "akmjb", T1, "mjld", "akbld")
    return sym_contract(W, "akbld", Uy, "klc", "abcd")


def _atrg_half_sym(T: SymTensor, chi: int) -> SymTensor:
    """_step!(::ATRG) on a Z_N tensor -- src/schemes/atrg.jl:47-82."""
    A, S, B, _ = sym_svd_trunc(T.permute((0, 2, 1, 3)), 2, chi)   # A [i1 i3 k], B [k i2 i4]
    Cc, D = sym_clone(A), sym_clone(B)
    Bs = sym_clone(B).scale_leg(0, S)
    Cc.scale_leg(2, S)
    # M[-1 -2;-3 -4] := B[-3;1 -4] * C[-1 1;-2]; SVD of permute(M, ((1,3),(2,4))) = [a c b d]
    M = sym_contract(Cc, "aib", Bs, "cid", "acbd")
    X, S2, Y, _ = sym_svd_trunc(M, 2, chi)
    rs = vec_map(S2, 1)
    X.scale_leg(2, rs)
    Y.scale_leg(0, rs)
    # Q[-1 -2;-3 -4] := A[3 -3;2] * D[1;-2 4] * X[4 2;-4] * Y[-1 1;3]
    AX = sym_contract(A, "kcj", X, "ljd", "kcld")
    YD = sym_contract(Y, "aik", D, "ibl", "akbl")
    Q = sym_contract(YD, "akbl", AX, "kcld", "abcd")
    H, S3, G, _ = sym_svd_trunc(Q, 2, chi)
    rs = vec_map(S3, 1)
    H.scale_leg(2, rs)
    G.scale_leg(0, rs)
    # T[-1 -2;-3 -4] := G[-1;-3 1] * H[1 -2;-4]
    return sym_contract(G, "aci", H, "ibd", "abcd")


def atrg_step_sym(T: SymTensor, chi: int) -> SymTensor:
    """step!(::ATRG) on a Z_N tensor -- src/schemes/atrg.jl:37-45."""
    T = _atrg_half_sym(T, chi).permute((1, 3, 0, 2))
    return _atrg_half_sym(T, chi).permute((2, 0, 3, 1))


# ------------------------------------------------------------------------------------
# 3D schemes on block-sparse tensors (HOTRG_3D / ATRG_3D on `classical_ising_3D(Z2Irrep)`,
# the tensor the reference's own 3D testsets run on: test/schemes.jl:8,365-383)
# ------------------------------------------------------------------------------------
NO_TRUNCATION = 10 ** 9   # truncrank large enough to keep every value (left_orth / right_orth)
LAST_PLAN = {}            # how the latest chunked step was split (for inspection / tests)


def sym_zeros(N, legs, ctx, flat=False):
    """All symmetry-allowed blocks, zero-initialised on the device.  `flat=True`: the blocks are
    views of ONE buffer, which is returned as well (one collective moves the whole tensor)."""
    import torch

    t = SymTensor(N, legs, {}, ctx)
    keys = list(t.keys())
    sizes = [max(1, math.prod(t.block_dims(k))) for k in keys]
    if flat:
        whole = torch.zeros(max(1, sum(sizes)), dtype=torch.float64, device=t.ctx.torch_device)
    off = 0
    for key, n in zip(keys, sizes):
        buf = whole[off: off + n] if flat else \
            torch.zeros(n, dtype=torch.float64, device=t.ctx.torch_device)
        off += n
        t.blocks[key] = DeviceTensor(buf, t.block_dims(key), None, t.ctx)
    return (t, whole) if flat else t


def leg_chunks(leg: Leg, size: int):
    """[(charge, lo, hi)] tiling every sector of `leg` into index ranges of at most `size`."""
    out = []
    for q in leg.charges:
        d = leg.dims[q]
        for lo in range(0, d, size):
            out.append((q, lo, min(d, lo + size)))
    return out


def sym_slice(T: SymTensor, axis: int, chunk) -> SymTensor:
    """T restricted to indices [lo, hi) of sector q of leg `axis` (compact copies of the blocks)."""
    q, lo, hi = chunk
    legs = list(T.legs)
    legs[axis] = Leg({q: hi - lo}, T.legs[axis].sign)
    out = SymTensor(T.N, legs, {}, T._ctx)
    for key, blk in T.blocks.items():
        if key[axis] != q:
            continue
        bd = list(T.block_dims(key))
        st = _colmajor_strides(bd)
        src = C.c_void_p(blk.buf.data_ptr() + 8 * lo * st[axis])
        bd[axis] = hi - lo
        piece = DeviceTensor.empty(bd, None, T.ctx)
        T.ctx.call("tnr_strided_copy", src, piece.ptr, len(bd), _lib.i64(bd), _lib.i64(st),
                   _lib.i64(_colmajor_strides(bd)))
        out.blocks[key] = piece
    return out


def sym_scatter(out: SymTensor, piece: SymTensor, where):
    """Writes `piece` (whose legs `axis` in `where` = {axis: (q, lo, hi)} are chunks) into the
    matching index ranges of the blocks of `out`."""
    for key, blk in piece.blocks.items():
        dst_blk = out.blocks[key]
        bd = piece.block_dims(key)
        dst_st = _colmajor_strides(out.block_dims(key))
        off = sum(lo * dst_st[ax] for ax, (_, lo, _) in where.items())
        dst = C.c_void_p(dst_blk.buf.data_ptr() + 8 * off)
        out.ctx.call("tnr_strided_copy", blk.ptr, dst, len(bd), _lib.i64(bd),
                     _lib.i64(_colmajor_strides(bd)), _lib.i64(dst_st))


def sym_trace_3d(T: SymTensor) -> float:
    """sum T[1 1; 2 3 2 3]  (finalize!(::HOTRG_3D / ::ATRG_3D), src/utility/finalize.jl:56-66)."""
    total = 0.0
    for key, blk in T.blocks.items():
        if key[0] != key[1] or key[2] != key[4] or key[3] != key[5]:
            continue
        d = T.block_dims(key)
        assert d[0] == d[1] and d[2] == d[4] and d[3] == d[5]
        st = _colmajor_strides(d)
        s = C.c_double()
        T.ctx.call("tnr_strided_sum", blk.ptr, 3, _lib.i64((d[0], d[2], d[3])),
                   _lib.i64((st[0] + st[1], st[2] + st[4], st[3] + st[5])), None, C.byref(s))
        total += s.value
    return total


def _hotrg3d_xproj_sym(A1: SymTensor, A2: SymTensor, chi: int) -> SymTensor:
    """_get_hotrg3d_xproj (hotrg3d.jl:87-100): eigh_trunc! of MMdag (open leg 6) and of MdagM
    (open leg 4), keep the one with the smaller truncation error.  Bosonic sectors: the twists
    of hotrg3d.jl:55-56,76-77 are identities."""
    A1c, A2c = sym_conj(A1), sym_conj(A2)
    # MM[x2 z z' x2'] := A2[z z2; Y2 X2 y2 x2] conj(A2'[z' z2; Y2 X2 y2 x2'])   (hotrg3d.jl:57-58)
    m2 = sym_contract(A2, "zabcdx", A2c, "wabcdy", "zxwy")
    m1 = sym_contract(A1, "azbcdx", A1c, "awbcdy", "zxwy")
    U, _, e = sym_eigh_trunc(sym_contract(m1, "zawc", m2, "zbwd", "abcd"), 2, chi)
    # MM[x2 z z' x2'] := conj(A2[z z2; Y2 x2 y2 X2]) A2'[z' z2; Y2 x2' y2 X2]   (hotrg3d.jl:78-79)
    m2 = sym_contract(A2c, "zabxcd", A2, "wabycd", "zxwy")
    m1 = sym_contract(A1c, "azbxcd", A1, "awbycd", "zxwy")
    U2, _, e2 = sym_eigh_trunc(sym_contract(m1, "zawc", m2, "zbwd", "abcd"), 2, chi)
    return U2 if e > e2 else U


def hotrg3d_substep_sym(T: SymTensor, chi: int, max_elems: int = 1 << 29, shard=None) -> SymTensor:
    """_step!(::HOTRG_3D) on a Z_N tensor -- src/schemes/hotrg3d.jl:102-129.

    The chi^8 intermediate of the pairwise contraction order is never formed beyond `max_elems`
    doubles: the two new open x-bonds (-4 and -6) are chunked, every (F, D) pair of chunks is
    contracted on its own and scattered into the index ranges [.., D, .., F] of the output
    blocks (same chunking as the dense engine, csrc/schemes.cu: hotrg3d_substep).

    `shard = (rank, world, group)`: the F chunks of the open x-bond -6 are dealt round-robin to
    the ranks (projectors and P_D are recomputed on every rank, as in the dense sharded step);
    every rank fills its index ranges of the zero-initialised output blocks, which live in one
    flat buffer, and one all-reduce (sum of disjoint supports: exact) replicates T'."""
    Ux = _hotrg3d_xproj_sym(T, T, chi)
    yperm = (0, 1, 3, 2, 5, 4)   # ((1,2),(4,3,6,5)), hotrg3d.jl:109
    Ty = T.permute(yperm)
    Uy = _hotrg3d_xproj_sym(Ty, Ty, chi)
    del Ty
    Uxc, Uyc = sym_conj(Ux), sym_conj(Uy)
    # T[-1 -2;-3 -4 -5 -6] := conj(Ux[x1 x2;-6]) Ux[x1' x2';-4] conj(Uy[y1 y2;-5]) Uy[y1' y2';-3]
    #                          A1[-1 z; y1' x1' y1 x1] A2[z -2; y2' x2' y2 x2]      (hotrg3d.jl:116-120)
    d = T.dims
    bx, by = Ux.legs[2], Uy.legs[2]
    out_legs = [T.legs[0], T.legs[1], by, bx, by.flipped(), bx.flipped()]
    nsect = T.N if T.N else max(1, len(bx.charges))
    base = d[0] * d[2] * d[4] * d[1] * d[2] * d[4] / max(1, nsect)   # R per unit |F||D|
    c = int(math.sqrt(max(1.0, max_elems / max(1.0, base))))
    rank, world, group = shard if shard is not None else (0, 1, None)
    if world > 1:
        # at least `world` chunks: shrink the chunk width until the x-bond splits that far
        c = min(c, max(bx.dims.values()))
        while c > 1 and len(leg_chunks(bx, c)) < world:
            c -= 1
    if world == 1 and c >= max(bx.dims.values()):
        chunks = [None]
    else:
        chunks = leg_chunks(bx, max(1, c))

    def Pd(D):
        u = Ux if D is None else sym_slice(Ux, 2, D)
        return sym_contract(T, "zbstuw", u, "qtd", "zbsuwqd")      # [z b y2' y2 x2 x1' d]

    cache = T.nnz() * bx.total <= 4 * max_elems     # all P_D together: nnz(T) * |x-bond| doubles
    mine = list(range(rank, len(chunks), world))
    LAST_PLAN["hotrg3d"] = {"chunks": len(chunks), "chunk_size": c, "cached_P": bool(cache),
                            "world": world, "my_F_chunks": len(mine)}
    P = [Pd(D) for D in chunks] if cache else None
    whole = None
    if chunks == [None]:
        out = None
    elif world > 1:
        out, whole = sym_zeros(T.N, out_legs, T.ctx, flat=True)
    else:
        out = sym_zeros(T.N, out_legs, T.ctx)
    for F in [chunks[i] for i in mine]:
        u = Uxc if F is None else sym_slice(Uxc, 2, F)
        Q = sym_contract(T, "azpqrx", u, "xwf", "azpqrwf")         # [a z y1' x1' y1 x2 f]
        for j, D in enumerate(chunks):
            R = sym_contract(Q, "azpqrwf", P[j] if cache else Pd(D), "zbsuwqd", "aprfbsud")
            R = sym_contract(R, "aprfbsud", Uyc, "rue", "apfbsde")
            R = sym_contract(R, "apfbsde", Uy, "psc", "abcdef")
            if out is None:
                return R
            sym_scatter(out, R, {3: D, 5: F})
    if whole is not None:
        import torch.distributed as dist

        dist.all_reduce(whole, op=dist.ReduceOp.SUM, group=group)
    return out


def hotrg3d_step_sym(T: SymTensor, chi: int, max_elems: int = 1 << 29, shard=None) -> SymTensor:
    """step!(::HOTRG_3D) on a Z_N tensor -- src/schemes/hotrg3d.jl:131-139."""
    for _ in range(3):
        T = hotrg3d_substep_sym(T, chi, max_elems, shard).permute((5, 3, 1, 2, 0, 4))  # ((6,4),(2,3,1,5))
    return T


def _orth_r(Z: SymTensor, ncod: int) -> SymTensor:
    """The factor `R` of `left_orth(Z)` up to the orthogonal gauge on its bond (which cancels
    in atrg3d.jl:58-66): Sigma V^T of the untruncated per-sector SVD.  Legs [r(+), dom...]."""
    _, S, Vt, _ = sym_svd_trunc(Z, ncod, NO_TRUNCATION)
    return Vt.scale_leg(0, S)


def _atrg3d_projectors_sym(Rl: SymTensor, Rr: SymTensor, chi: int):
    """atrg3d.jl:58-66: temp = Rl Rr; U S V = svd_trunc(temp);
    Pa[p q; k] = Rr V' S^-1/2,  Pb[k; p q] = S^-1/2 U' Rl."""
    t = sym_contract(Rl, "ipq", Rr, "pqj", "ij")
    U, S, V, _ = sym_svd_trunc(t, 1, chi)
    inv = vec_map(S, 2, -0.5)
    Pa = sym_contract(Rr, "pqj", sym_conj(V), "kj", "pqk").scale_leg(2, inv)
    Pb = sym_contract(sym_conj(U), "ik", Rl, "ipq", "kpq").scale_leg(0, inv)
    return Pa, Pb


def _split_for(leg: Leg, world: int, size=None):
    """Chunks of `leg` for `world` ranks: the widest chunk width (<= `size`) that still yields at
    least `world` chunks (fewer only when the leg has fewer than `world` states)."""
    c = max(leg.dims.values())
    if size is not None:
        c = max(1, min(c, int(size)))
    while c > 1 and len(leg_chunks(leg, c)) < world:
        c -= 1
    return leg_chunks(leg, c)


def _tsqr_orth_r(pieces, row_legs, col_legs, N, ctx, nslots, shard):
    """R factor of a tall operand given as row chunks (TSQR): `pieces` = [(slot, R_slot)] are the
    R factors [r(+); cols] of this rank's chunks; they are stacked along the bond -- slot s owns
    rows [s n_c, (s+1) n_c) of coupled sector c (n_c = columns of that sector; unused rows stay
    zero and drop out of R^T R) -- in ONE flat zero-initialised buffer, one all-reduce (sum of
    disjoint supports: exact) replicates the stack, and its own R factor is the result.  Same
    orthogonal gauge freedom as `_orth_r`.  Only coupled sectors that the full operand has
    (`row_legs` x `col_legs`) enter, as in `SymTensor.matricize`."""
    rank, world, group = shard
    probe = SymTensor(N, list(row_legs) + list(col_legs), {}, ctx)
    nrow = len(row_legs)
    rt = probe._tuples(range(nrow), False)
    ct = probe._tuples(range(nrow, nrow + len(col_legs)), True)
    ncols = {c: v[-1][1] + v[-1][2] for c, v in ct.items() if c in rt}
    stack_leg = Leg({c: nslots * n for c, n in ncols.items()}, +1)
    stack, whole = sym_zeros(N, [stack_leg] + list(col_legs), ctx, flat=True)
    for slot, R in pieces:
        for key, blk in R.blocks.items():
            dst_blk = stack.blocks[key]
            bd = R.block_dims(key)
            assert bd[0] <= ncols[key[0]]
            dst = C.c_void_p(dst_blk.buf.data_ptr() + 8 * slot * ncols[key[0]])
            ctx.call("tnr_strided_copy", blk.ptr, dst, len(bd), _lib.i64(bd),
                     _lib.i64(_colmajor_strides(bd)),
                     _lib.i64(_colmajor_strides(stack.block_dims(key))))
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(whole, op=dist.ReduceOp.SUM, group=group)
    return _orth_r(stack, 1)


def _atrg3d_tail_sharded(gU, fU, fV, gV, chi: int, shard, chunk=None) -> SymTensor:
    """atrg3d.jl:47-82 with the open bond -1 of AX and YD cut into chunks that are dealt
    round-robin to the ranks `shard = (rank, world, group)`.  No rank forms a whole chi^6 tensor:
    a chunk of AX / YD is contracted, reduced to its two R factors (rows of the tall
    matricizations: TSQR, `_tsqr_orth_r`), and -- once the replicated projectors exist --
    projected into its index range of H / G.  Exchanged: four R stacks (chi^2 x chi^2 per chunk)
    and H, G (chi^4), each by one all-reduce of a flat block buffer with disjoint supports.
    The two truncated SVDs in front of it (atrg3d.jl:35-45) stay replicated."""
    rank, world, group = shard
    N, ctx = gU.N, gU.ctx
    q = (0, 1, 4, 5, 2, 3)

    def passes(left, right, chunks):
        # operand[-1 -2;-3 -4 -5 -6] := left[1 -2;-3 -5] right[-1 1;-4 -6], chunked along -1
        mine = [(s, ch) for s, ch in enumerate(chunks) if s % world == rank]
        legs = [left.legs[3], right.legs[0], right.legs[1], left.legs[1], right.legs[2], left.legs[2]]
        parts, Ra, Rb = [], [], []
        for s, ch in mine:
            Z = sym_contract(sym_slice(left, 3, ch), "idfa", right, "bcei", "abcdef")
            Ra.append((s, _orth_r(Z, 4)))                    # [r; 5 6]
            Rb.append((s, _orth_r(Z.permute(q), 4)))         # [r; 3 4]
            parts.append((ch, Z))
        n = len(chunks)
        R56 = _tsqr_orth_r(Ra, legs[:4], legs[4:], N, ctx, n, shard)
        R34 = _tsqr_orth_r(Rb, legs[:2] + legs[4:], legs[2:4], N, ctx, n, shard)
        return parts, legs, R56, R34

    cA = _split_for(gU.legs[3], world, chunk)
    cY = _split_for(fV.legs[3], world, chunk)
    LAST_PLAN["atrg3d"] = {"world": world, "chunks_AX": len(cA), "chunks_YD": len(cY),
                           "my_chunks": sum(1 for s in range(len(cA)) if s % world == rank)
                           + sum(1 for s in range(len(cY)) if s % world == rank)}
    # AX[-1 -2;-3 -4 -5 -6] := A[1 -2;-3 -5] X[-1 1;-4 -6];  YD := Y[1 -2;-3 -5] D[-1 1;-4 -6]
    ax_parts, ax_legs, R2t, R4t = passes(gU, fU, cA)
    yd_parts, yd_legs, R1, R3 = passes(fV, gV, cY)
    P1, P2 = _atrg3d_projectors_sym(R1, R2t.permute((1, 2, 0)), chi)
    P3, P4 = _atrg3d_projectors_sym(R3, R4t.permute((1, 2, 0)), chi)

    def project(parts, legs, Pc, lc, Pd, ld):
        out, whole = sym_zeros(N, [legs[0], legs[1], Pc.legs[lc.index("c")], Pd.legs[ld.index("d")]],
                               ctx, flat=True)
        for ch, Z in parts:
            piece = sym_contract(sym_contract(Z, "abijkl", Pc, lc, "abklc"), "abklc", Pd, ld, "abcd")
            sym_scatter(out, piece, {0: ch})
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(whole, op=dist.ReduceOp.SUM, group=group)
        return out

    # H[-1 -2;-3 -4] := YD[-1 -2;1 2 3 4] Proj_3[1 2;-3] Proj_1[3 4;-4]
    H = project(yd_parts, yd_legs, P3, "ijc", P1, "kld")
    del yd_parts
    # G[-1 -2;-3 -4] := AX[-1 -2;1 2 3 4] Proj_4[-3;1 2] Proj_2[-4;3 4]
    G = project(ax_parts, ax_legs, P4, "cij", P2, "dkl")
    del ax_parts
    # T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] H[-1 1;-3 -4]
    return sym_contract(H, "aicd", G, "ibef", "abcdef")


def atrg3d_substep_sym(T: SymTensor, chi: int, shard=None, chunk=None) -> SymTensor:
    """_step!(::ATRG_3D) on a Z_N tensor -- src/schemes/atrg3d.jl:34-83.  `permute(X, ((4,1),(2,3)))`
    of the reference is expressed through leg labels instead of data movement, as in the dense
    engine (csrc/schemes.cu: atrg3d_substep).

    `shard = (rank, world, group)` (or `chunk` = a chunk width on one process): AX / YD are
    never formed whole, their open bond -1 is chunked and the chunks are dealt to the ranks
    (`_atrg3d_tail_sharded`)."""
    perm = (1, 4, 5, 2, 3, 0)   # ((2,5,6),(3,4,1))
    fU, fS, fV, _ = sym_svd_trunc(T.permute(perm), 3, chi)     # U [i2 i5 i6 k], V [k i3 i4 i1]
    US = sym_clone(fU).scale_leg(3, fS)
    SV = sym_clone(fV).scale_leg(0, fS)
    # M[-1 -2;-3 -4 -5 -6] := B[1 -2;-3 -4] C[-1 1;-5 -6], produced directly as permute(M, perm)
    Mp = sym_contract(US, "iefa", SV, "bcdi", "befcda")
    del US, SV
    gU, gS, gV, _ = sym_svd_trunc(Mp, 3, chi)                   # U [m2 m5 m6 k], V [k m3 m4 m1]
    del Mp
    rs = vec_map(gS, 1)
    gU.scale_leg(3, rs)    # X
    gV.scale_leg(0, rs)    # Y
    if (shard is not None and shard[1] > 1) or chunk is not None:
        return _atrg3d_tail_sharded(gU, fU, fV, gV, chi, shard if shard is not None
                                    else (0, 1, None), chunk)
    # AX[-1 -2;-3 -4 -5 -6] := A[1 -2;-3 -5] X[-1 1;-4 -6];  YD := Y[1 -2;-3 -5] D[-1 1;-4 -6]
    AX = sym_contract(gU, "idfa", fU, "bcei", "abcdef")
    YD = sym_contract(fV, "idfa", gV, "bcei", "abcdef")
    q = (0, 1, 4, 5, 2, 3)
    R1 = _orth_r(YD, 4)                                  # [r; 5 6]
    R2 = _orth_r(AX, 4).permute((1, 2, 0))               # [5 6; r]
    R3 = _orth_r(YD.permute(q), 4)                       # [r; 3 4]
    R4 = _orth_r(AX.permute(q), 4).permute((1, 2, 0))    # [3 4; r]
    P1, P2 = _atrg3d_projectors_sym(R1, R2, chi)
    P3, P4 = _atrg3d_projectors_sym(R3, R4, chi)
    # H[-1 -2;-3 -4] := YD[-1 -2;1 2 3 4] Proj_3[1 2;-3] Proj_1[3 4;-4]
    H = sym_contract(sym_contract(YD, "abijkl", P3, "ijc", "abklc"), "abklc", P1, "kld", "abcd")
    # G[-1 -2;-3 -4] := AX[-1 -2;1 2 3 4] Proj_4[-3;1 2] Proj_2[-4;3 4]
    G = sym_contract(sym_contract(AX, "abijkl", P4, "cij", "abklc"), "abklc", P2, "dkl", "abcd")
    # T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] H[-1 1;-3 -4]
    return sym_contract(H, "aicd", G, "ibef", "abcdef")


def atrg3d_step_sym(T: SymTensor, chi: int, shard=None, chunk=None) -> SymTensor:
    """step!(::ATRG_3D) on a Z_N tensor -- src/schemes/atrg3d.jl:85-97."""
    for _ in range(3):
        T = atrg3d_substep_sym(T, chi, shard, chunk).permute((3, 5, 1, 4, 0, 2))   # ((4,6),(2,5,1,3))
    return T
