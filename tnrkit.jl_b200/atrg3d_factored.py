"""ATRG_3D without any chi^6 object: the `_step!` of /root/reference/src/schemes/atrg3d.jl:34-83
on a tensor that is kept in the factored form in which the step produces it.

The last line of `_step!` is `T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] * H[-1 1;-3 -4]`: a chi^6
tensor (98 GB at chi = 48) that is the product of two chi^4 tensors over one bond.  Every later
use of it is linear, so it never has to exist:

  * the two truncated SVDs (atrg3d.jl:35, 43) keep chi of chi^3 singular triplets.  They run as a
    block subspace iteration with Rayleigh-Ritz on the implicit operator -- one product with the
    matricized tensor is two contractions of cost chi^5 b through the bond instead of one of
    cost chi^6 b -- certified by the residuals ||A v - sigma u|| <= tol sigma_1 of the kept
    triplets (same iteration and acceptance rule as `svd_topk` in csrc/tensor_ops.cu, which
    does it for an explicit matrix).  `M = B*C` (atrg3d.jl:41) is again a two-factor object.
  * `AX`, `YD` (atrg3d.jl:49-50) are two-factor objects as well.  Their R factors
    (atrg3d.jl:53-56) need the tall chi^4 x chi^2 matricizations, which are produced CHUNK BY
    CHUNK along an open bond (the new bond `a` of X for AX, `b` of Y for YD): every chunk gives
    an R factor (`tnr_orth_r`), and the R factor of the stack of chunk factors is the R factor of
    the whole (TSQR).  The same chunks are formed a second time for `H` / `G` (atrg3d.jl:68-69).

Multi-GPU (`shard=True` under torch.distributed, one process per GPU): the chunks of the open
bond are divided between the ranks.  The exchanges are one all-gather of the stacked chunk R
factors (chi^2 x chi^2 each) per matricization and one all-gather of `H` / `G` along the open
bond (their last leg, contiguous slabs); the subspace iterations (O(chi^6) flop) are replicated.

Cost per `_step!`: O(chi^7) for the chunks + 4 x 2 chi^8 for the R factors, memory O(chi^5 c) for
chunks of width c, instead of O(chi^9) / O(chi^6).

This file only sequences C-ABI calls (`tnr_contract`, `tnr_svd_trunc`, `tnr_orth_r`,
`tnr_axis_scale`, `tnr_vec_map`, `tnr_gemm`, `tnr_permute`, `tnr_scale`); all arithmetic runs in
libtnrcuda.
"""
from __future__ import annotations

import logging
import math

import numpy as np

from . import _lib
from .tensor import DeviceTensor, contract, svd_trunc

log = logging.getLogger("tnrkit.jl_b200")

NO_TRUNCATION = 1 << 30
LEGS = "abcdef"
BOND = "i"
BLOCK = "k"          # label of the block index of the subspace iteration
LAST_STATS = {}      # filled by the last substep: iteration counts, chunk plan (for tests / bench)
PROFILE = False      # True: synchronise after every phase of a substep and record its seconds


class _Phases:
    """Wall-clock seconds per phase of a substep (only with PROFILE: it synchronises the stream)."""

    def __init__(self, ctx):
        import time

        self.ctx, self.t, self.out, self.clock = ctx, None, {}, time.perf_counter
        if PROFILE:
            ctx.synchronize()
            self.t = self.clock()

    def mark(self, name):
        if self.t is not None:
            self.ctx.synchronize()
            now = self.clock()
            self.out[name] = round(self.out.get(name, 0.0) + now - self.t, 6)
            self.t = now


def _view(t: DeviceTensor, dims) -> DeviceTensor:
    assert math.prod(dims) == t.size, (dims, t.dims)
    return DeviceTensor(t.buf, dims, None, t.ctx)


def _slice_last(t: DeviceTensor, lo: int, hi: int) -> DeviceTensor:
    """t[..., lo:hi] -- contiguous in the column-major layout, so a view."""
    slab = math.prod(t.dims[:-1])
    return DeviceTensor(t.buf[lo * slab: hi * slab], t.dims[:-1] + (hi - lo,), None, t.ctx)


def _scale_leg(t: DeviceTensor, axis: int, s: DeviceTensor, mode: int = 0, p: float = 0.0):
    """t[..., i_axis, ...] *= f(s[i_axis]) in place (f: identity / sqrt / pseudopow)."""
    m1 = math.prod(t.dims[:axis])
    m2 = math.prod(t.dims[axis + 1:])
    t.ctx.call("tnr_axis_scale", t.ptr, m1, t.dims[axis], m2, s.ptr, mode, float(p))
    return t


def _vec_map(s: DeviceTensor, mode: int, p: float = 0.0) -> DeviceTensor:
    o = DeviceTensor.empty(s.dims, 1, s.ctx)
    s.ctx.call("tnr_vec_map", s.ptr, o.ptr, s.size, mode, float(p))
    return o


class TwoFactor:
    """T[legs] = sum_i P[lp] Q[lq]: a 6-leg tensor held as two factors joined by the bond `i`.
    `legs` is the current leg order (a permutation of the open labels of lp and lq), so
    `permute` costs nothing; the label of a leg is only a name."""

    def __init__(self, P: DeviceTensor, lp: str, Q: DeviceTensor, lq: str, legs: str):
        assert len(lp) == len(P.dims) and len(lq) == len(Q.dims)
        assert [c for c in lp if c in lq] == [BOND], (lp, lq)
        assert sorted(legs) == sorted((lp + lq).replace(BOND, "")), (lp, lq, legs)
        self.P, self.lp, self.Q, self.lq, self.legs = P, lp, Q, lq, legs
        self.ctx = P.ctx

    # -- bookkeeping --------------------------------------------------------------------
    def dim(self, label: str) -> int:
        return self.P.dims[self.lp.index(label)] if label in self.lp else \
            self.Q.dims[self.lq.index(label)]

    @property
    def dims(self):
        return tuple(self.dim(c) for c in self.legs)

    @property
    def bond_dim(self):
        return self.dim(BOND)

    def permute(self, perm):
        """TensorKit.permute: new leg k is old leg perm[k]; only the names move."""
        return TwoFactor(self.P, self.lp, self.Q, self.lq, "".join(self.legs[p] for p in perm))

    def relabel(self):
        """Same tensor with its legs named a..f in their current order."""
        ren = {old: new for old, new in zip(self.legs, LEGS[: len(self.legs)])}
        ren[BOND] = BOND
        tr = lambda s: "".join(ren[c] for c in s)  # noqa: E731
        return TwoFactor(self.P, tr(self.lp), self.Q, tr(self.lq), LEGS[: len(self.legs)])

    @classmethod
    def from_dense(cls, T: DeviceTensor):
        """Exact factorization T[a b c d e f] = sum_i P[a c d i] Q[b e f i] of an explicit tensor
        (the model tensor handed to the constructor): untruncated SVD of the (a c d) x (b e f)
        matricization, the split in which the step itself leaves its result."""
        assert len(T.dims) == 6
        Tp = T.permute((0, 2, 3, 1, 4, 5))
        U, S, Vt, _ = svd_trunc(Tp, 3, NO_TRUNCATION)
        _scale_leg(U, 3, S)
        Q = Vt.permute((1, 2, 3, 0))
        return cls(U, "acdi", Q, "befi", LEGS)

    def to_dense(self) -> DeviceTensor:
        """The explicit tensor (prod(dims) doubles): for inspection and for small cases."""
        out = contract(self.P, self.lp, self.Q, self.lq, self.legs)
        out.ncod = 2
        return out

    def scale(self, alpha: float):
        self.ctx.call("tnr_scale", self.P.ptr, self.P.size, float(alpha))
        return self

    # -- the linear maps the step needs -------------------------------------------------
    def apply(self, rows: str, cols: str, V: DeviceTensor) -> DeviceTensor:
        """out[rows..., k] = sum_cols T[rows; cols] V[cols..., k], through the bond: the factor
        that holds more of the summed legs is contracted first, so no intermediate is larger
        than (open legs of that factor) x k."""
        lv = cols + BLOCK
        np_, nq_ = sum(c in self.lp for c in cols), sum(c in self.lq for c in cols)
        (F1, l1), (F2, l2) = ((self.P, self.lp), (self.Q, self.lq)) if np_ >= nq_ else \
            ((self.Q, self.lq), (self.P, self.lp))
        lw = "".join(c for c in l1 if c not in cols) + "".join(c for c in cols if c not in l1) + BLOCK
        W = contract(F1, l1, V, lv, lw)
        return contract(F2, l2, W, lw, rows + BLOCK)

    def trace_3d(self) -> float:
        """T[1 1; 2 3 2 3] (src/utility/finalize.jl:56-66) in the current leg order."""
        L = self.legs
        ren = {L[1]: L[0], L[4]: L[2], L[5]: L[3]}
        lp = "".join(ren.get(c, c) for c in self.lp)
        lq = "".join(ren.get(c, c) for c in self.lq)
        if len(set(lp)) < len(lp) or len(set(lq)) < len(lq):
            # both legs of a traced pair sit in one factor: not the split the step produces
            raise NotImplementedError("trace of a TwoFactor whose traced pairs are not split "
                                      "between the factors")
        return float(contract(self.P, lp, self.Q, lq, "").to_numpy().reshape(-1)[0])

    def __repr__(self):
        return (f"TwoFactor(dims={self.dims}, bond={self.bond_dim}: "
                f"P[{self.lp}] {self.P.dims} * Q[{self.lq}] {self.Q.dims})")


# ---------------------------------------------------------------------------------------
# truncated SVD of an implicit operator
# ---------------------------------------------------------------------------------------
def _residual_norms(Z: DeviceTensor, Ul: DeviceTensor, sig: DeviceTensor, m: int, k: int,
                    eye: DeviceTensor):
    """|| Z[:, j] - sig[j] Ul[:, j] ||_2 for j < k (Z = A V, Ul = left vectors), on the device;
    only the k norms^2 travel to the host.  `eye`: k x k identity (D -= E I through tnr_gemm)."""
    ctx = Z.ctx
    D = DeviceTensor(Z.buf[: m * k].clone(), (m, k), None, ctx)
    E = DeviceTensor(Ul.buf[: m * k].clone(), (m, k), None, ctx)
    _scale_leg(E, 1, sig)
    ctx.call("tnr_gemm", b"N", b"N", m, k, k, -1.0, E.ptr, m, eye.ptr, k, 1.0, D.ptr, m)
    g = contract(D, "mj", D, "ml", "jl").to_numpy()
    return np.sqrt(np.maximum(np.diag(g), 0.0))


def _orthonormalize(M: DeviceTensor) -> bool:
    """M (rows x cols, tall) <- orthonormal basis of its columns, in place (`tnr_orthonormalize`:
    CholeskyQR2).  False: refused (ill-conditioned or rank deficient), M untouched."""
    import ctypes as C

    rows, cols = M.dims
    if rows < cols:
        return False
    refused = C.c_int(1)
    M.ctx.call("tnr_orthonormalize", M.ptr, rows, cols, C.byref(refused))
    return refused.value == 0


# (rows, cols, block, kept) of an operator -> (iterations, geometric rate) of the last certified
# subspace iteration on an operator of that shape: successive RG steps of a run converge alike, so
# the first Rayleigh-Ritz check of the next one is placed just before the expected end instead of
# at iterations 2 and 4 (a wrong hint only moves a check; acceptance is by the residual alone)
_CHECK_HINT = {}


def _next_check(history, it: int, rel: float, tol: float, max_jump: int = 8,
                rate_hint: float | None = None) -> int:
    """Iteration of the next Rayleigh-Ritz check: where the geometric rate seen between the last
    two checks (or `rate_hint`, before two checks exist) reaches tol / 2, at most `max_jump`
    iterations ahead; two iterations ahead while no rate is known or the residual did not
    shrink."""
    rate = None
    if history:
        it0, rel0 = history[-1]
        if 0.0 < rel < rel0 and it > it0:
            rate = (rel / rel0) ** (1.0 / (it - it0))
    elif rate_hint is not None and 0.0 < rate_hint < 1.0:
        rate = rate_hint
    if rate is not None and rel > 0.0:
        need = math.log(0.5 * tol / rel) / math.log(rate)
        return it + int(min(max_jump, max(1, math.ceil(need))))
    return it + 2


def _apply_sharded(F: TwoFactor, rows: str, cols: str, V: DeviceTensor, shard) -> DeviceTensor:
    """F.apply(rows, cols, V) with the block columns of V (its last leg) dealt to the ranks of a
    sharded run and the products all-gathered along that leg: every rank ends with the same
    bytes, so everything computed from them stays bit-identical between the replicas."""
    if shard is None or shard[3] == 1 or V.dims[-1] < shard[3]:
        return F.apply(rows, cols, V)
    from .schemes import allgather_last_leg, shard_range

    _, group, rank, world = shard
    nb = V.dims[-1]
    lo, hi = shard_range(nb, rank, world)
    out = DeviceTensor.empty(tuple(F.dim(c) for c in rows) + (nb,), None, F.ctx)
    if hi > lo:
        part = F.apply(rows, cols, _slice_last(V, lo, hi))
        _slice_last(out, lo, hi).buf.copy_(part.buf[: part.size])
    allgather_last_leg(out.buf, out.dims, group)
    return out


def svd_topk_factored(F: TwoFactor, rows: str, cols: str, chi: int, tol: float = 1e-13,
                      maxit: int = 400, seed: int = 0x5EED, stats: dict | None = None,
                      block: int | None = None, dense_fallback_elems: int = 1 << 27,
                      cholqr: bool = True, shard=None):
    """svd_trunc(permute(T, (rows), (cols)); trunc = truncrank(chi)) for a TwoFactor T, without
    forming T.  Returns U [rows..., k], S [k], V [cols..., k]  (V is the TRANSPOSE of TensorKit's
    third factor: callers address legs by label, so no data is moved to transpose it).

    Block subspace iteration on the right singular subspace, block b = max(2 chi, chi + 64):
    Z = A Q,  U = orth(Z),  Y = A^T U,  Q = orth(Y), with CholeskyQR2 (`tnr_orthonormalize`) as
    `orth`; at the check iterations (placed by `_next_check`) the Rayleigh-Ritz step instead:
    Y = V' S' X^T (thin SVD)  =>  A ~ (U X) S' V'^T, accepted when
    max_j<chi ||A v_j - s_j u_j|| <= tol s_1 (or stalled below 20 tol, the rounding floor of the
    products).  Where CholeskyQR2 refuses (rank-deficient or ill-conditioned block: the first RG
    steps) the thin SVD Z = U S W^T takes its place, as in round 1 (`cholqr=False`: always).
    Exact (one dense SVD) when b reaches min(rows, cols).
    An iteration that does not certify falls back to the dense SVD of the materialised matrix
    when that has at most `dense_fallback_elems` entries, and raises otherwise: the result
    never depends on an uncertified subspace."""
    ctx = F.ctx
    rd = tuple(F.dim(c) for c in rows)
    cd = tuple(F.dim(c) for c in cols)
    m, n = math.prod(rd), math.prod(cd)
    r = min(m, n)
    k = min(chi, r)
    b = max(2 * k, k + 64) if block is None else max(int(block), k)
    st = stats if stats is not None else {}

    def dense_svd(why):
        dense = contract(F.P, F.lp, F.Q, F.lq, rows + cols)
        U, S, Vt, _ = svd_trunc(dense, len(rows), chi)
        st.update(dense=True, block=r, why=why)
        return U, S, Vt.permute(tuple(range(1, len(cols) + 1)) + (0,))

    if b >= r:
        st.update(iterations=0)
        return dense_svd("block covers the matrix")
    Q = DeviceTensor.empty(cd + (b,), None, ctx)      # start block: generated on the device
    ctx.call("tnr_fill_random", Q.ptr, Q.size, int(seed))
    ph = _Phases(ctx)

    def op(r_, c_, V):      # operator products; block columns dealt to the ranks of a sharded run
        return _apply_sharded(F, r_, c_, V, shard)

    Z = op(rows, cols, Q)
    ph.mark("apply")
    best, stalled = math.inf, 0
    Ul = sig = None
    eyes = {}
    # Rayleigh-Ritz (thin SVD of A^T U, residuals) only at the iterations `check`; in between
    # the bases are just re-orthonormalised (CholeskyQR2 on the tensor cores).  The first checks
    # give the convergence rate, later ones are placed where the tolerance is predicted.
    hint_key = (m, n, b, k, tol)
    hint = _CHECK_HINT.get(hint_key) if cholqr else None
    check, history, checks, cheap_its = min(max(2, hint[0] - 1) if hint else 2, maxit), [], 0, 0
    it = 0
    while True:
        it += 1
        bz = Z.dims[-1]
        Zm = _view(Z, (m, bz))
        if cholqr and _orthonormalize(Zm):                 # U = orth(Z) in place
            U, keep = Z, bz
            ph.mark("orth")
        else:
            U, S, _, _ = svd_trunc(Zm, 1, NO_TRUNCATION)                          # Z = U S W^T
            s = S.to_numpy()
            ph.mark("svd_Z")
            keep = int(np.count_nonzero(s > 1e-14 * s[0]))
            if keep == 0:
                return dense_svd("zero operator")
            U = DeviceTensor(U.buf[: m * keep], rd + (keep,), None, ctx)
        ke = min(k, keep)     # rank(A) < chi: the block spans the whole range, the rest is zero
        Y = op(cols, rows, U)                                               # A^T U
        ph.mark("apply")
        if cholqr and it < check and _orthonormalize(_view(Y, (n, keep))):       # Q = orth(A^T U)
            ph.mark("orth")
            Z = op(rows, cols, Y)
            ph.mark("apply")
            cheap_its += 1
            continue
        checks += 1
        if ke not in eyes:
            eyes[ke] = DeviceTensor.from_numpy(np.eye(ke), None, ctx)
        Vh, sig, Xt, _ = svd_trunc(_view(Y, (n, keep)), 1, NO_TRUNCATION)        # Y = Vh sig Xt
        ph.mark("svd_Y")
        Ul = contract(_view(U, (m, keep)), "mj", Xt, "lj", "ml")                 # left vectors
        Q = _view(Vh, cd + (keep,))
        ph.mark("rotate")
        Z = op(rows, cols, Q)                                               # A Vh (next Z)
        ph.mark("apply")
        res = _residual_norms(_view(Z, (m, keep)), Ul, sig, m, ke, eyes[ke])
        ph.mark("residual")
        if ph.out:
            st["phase_s"] = dict(ph.out)
        smax = float(sig.to_numpy()[0])
        rel = float(res.max()) / smax if smax > 0.0 else 0.0
        if not math.isfinite(rel):
            raise _lib.TNRCudaError("svd_topk_factored: non-finite residual")
        if rel <= tol:
            break
        # slow but steady convergence (flat 3D spectra: rate (s_{b+1}/s_chi)^2 per iteration) is
        # not a stall; only < 10 % progress over the best residual counts
        stalled = stalled + 1 if rel > 0.9 * best else 0
        best = min(best, rel)
        if stalled >= 3 and best <= 20 * tol:
            break
        if stalled >= 12 or it >= maxit:
            if m * n <= dense_fallback_elems:
                log.warning("svd_topk_factored: residual %.2e after %d iterations (block %d); "
                            "dense SVD of the %d x %d matrix instead", best, it, b, m, n)
                st.update(iterations=it)
                return dense_svd("subspace iteration did not certify")
            raise _lib.TNRCudaError(f"svd_topk_factored: no convergence (residual {best:.2e} "
                                    f"after {it} iterations, block {b})")
        check = min(maxit, _next_check(history, it, rel, tol, rate_hint=hint[1] if hint else None))
        history.append((it, rel))
    st.update(checks=checks, cheap_iterations=cheap_its)
    if it >= 4 and rel > 0.0:
        # overall geometric rate of this run, from the size of a random start (~1) to `rel`
        _CHECK_HINT[hint_key] = (it, min(0.95, max(1e-3, rel ** (1.0 / it))))
    st.update(iterations=it, dense=False, block=b, residual=rel, rank=ke)
    if ke == k:
        Uk = DeviceTensor(Ul.buf[: m * k].clone(), rd + (k,), None, ctx)
        Sk = DeviceTensor(sig.buf[:k].clone(), (k,), None, ctx)
        Vk = DeviceTensor(Q.buf[: n * k].clone(), cd + (k,), None, ctx)
        return Uk, Sk, Vk
    # fewer than chi nonzero singular values: the remaining triplets have sigma = 0 and enter
    # every later contraction with weight sigma or sqrt(sigma); they are stored as zeros
    import torch

    def padded(src, lead):
        buf = torch.zeros(lead * k, dtype=torch.float64, device=ctx.torch_device)
        buf[: lead * ke].copy_(src.buf[: lead * ke])
        return buf

    return (DeviceTensor(padded(Ul, m), rd + (k,), None, ctx),
            DeviceTensor(padded(sig, 1), (k,), None, ctx),
            DeviceTensor(padded(Q, n), cd + (k,), None, ctx))


# ---------------------------------------------------------------------------------------
# chunked / sharded pieces
# ---------------------------------------------------------------------------------------
def _dist(shard, group):
    if not shard:
        return None, 0, 1
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


def chunk_plan(n: int, rank: int, world: int, width: int):
    """Chunks [lo, hi) of an open bond of dimension n owned by `rank`: the bond is divided into
    `world` contiguous blocks (schemes.shard_range), each block into pieces of at most `width`."""
    from .schemes import shard_range

    lo, hi = shard_range(n, rank, world)
    return [(c, min(c + width, hi)) for c in range(lo, hi, width)]


def _orth_r(D: DeviceTensor, ncod: int) -> DeviceTensor:
    """R factor (up to the orthogonal gauge) of the matricization of D with `ncod` row legs,
    as a matrix [r, cols]."""
    m, n = math.prod(D.dims[:ncod]), math.prod(D.dims[ncod:])
    if m >= n:
        R = DeviceTensor.empty((n, n), 1, D.ctx)
        D.ctx.call("tnr_orth_r", D.ptr, len(D.dims), _lib.i64(D.dims), ncod, R.ptr)
        return R
    _, S, Vt, _ = svd_trunc(D, ncod, NO_TRUNCATION)      # wide chunk (tiny cases): R = S V^T, m x n
    Vt = _view(Vt, (m, n))
    return _scale_leg(Vt, 0, S)


class _ChunkedPair:
    """A two-factor tensor Z[a b c d e f] whose dense form is produced in chunks of one open
    bond that is the LAST leg of one factor (so a chunk of the factor is a view)."""

    def __init__(self, F: TwoFactor, chunk_label: str, width: int, rank: int, world: int):
        self.F, self.label = F, chunk_label
        if F.lp[-1] == chunk_label:
            self.in_p = True
        elif F.lq[-1] == chunk_label:
            self.in_p = False
        else:
            raise ValueError("chunk leg must be the last leg of a factor")
        n = F.dim(chunk_label)
        self.n, self.world, self.rank = n, world, rank
        self.plans = [chunk_plan(n, r, world, width) for r in range(world)]

    def dense(self, lo, hi) -> DeviceTensor:
        F = self.F
        if self.in_p:
            return contract(_slice_last(F.P, lo, hi), F.lp, F.Q, F.lq, F.legs)
        return contract(F.P, F.lp, _slice_last(F.Q, lo, hi), F.lq, F.legs)


def _stack_r(pieces, n: int, plans_rows, rank: int, world: int, dist, group, ctx):
    """R factor of the row-stack of all chunk R factors of all ranks.  The stack is held
    transposed, W = [R_1^T R_2^T ...] (n x K, every rank's columns one contiguous slab, padded
    with zero columns -- zero rows of the stack -- to the same width), all-gathered, and
    R = Sigma U^T from W = U Sigma V^T."""
    if world == 1 and len(pieces) == 1:
        return pieces[0]
    per_rank = [sum(rows) for rows in plans_rows]
    width = max(per_rank)
    import torch

    W = DeviceTensor(torch.zeros(n * width * world, dtype=torch.float64, device=ctx.torch_device),
                     (n, width * world), None, ctx)
    off = rank * width
    for R in pieces:
        kk = R.dims[0]
        dst = DeviceTensor(W.buf[off * n: (off + kk) * n], (n, kk), None, ctx)
        ctx.call("tnr_permute", R.ptr, dst.ptr, 2, _lib.i64(R.dims), _lib.i32((1, 0)))
        off += kk
    if world > 1:
        from .schemes import allgather_last_leg

        allgather_last_leg(W.buf, W.dims, group)
    U, S, _, _ = svd_trunc(W, 1, NO_TRUNCATION)          # U: n x k
    R = U.permute((1, 0))
    return _scale_leg(R, 0, S)


def _r_factors(Z: _ChunkedPair, dist, group):
    """(R_ef, R_cd): R factors of Z as (a b c d) x (e f) and as (a b e f) x (c d)
    (left_orth of YD / right_orth of AX, atrg3d.jl:53-56), each [r, pair]."""
    ctx = Z.F.ctx
    d = {c: Z.F.dim(c) for c in LEGS}
    n_ef, n_cd = d["e"] * d["f"], d["c"] * d["d"]
    mine_ef, mine_cd = [], []
    for lo, hi in Z.plans[Z.rank]:
        D = Z.dense(lo, hi)
        mine_ef.append(_orth_r(D, 4))
        Dq = D.permute((0, 1, 4, 5, 2, 3))
        del D
        mine_cd.append(_orth_r(Dq, 4))
        del Dq

    def rows(n):
        out = []
        for plan in Z.plans:
            rr = []
            for lo, hi in plan:
                full = {**d, Z.label: hi - lo}
                rr.append(min(n, math.prod(full[c] for c in LEGS) // n))
            out.append(rr)
        return out

    R_ef = _stack_r(mine_ef, n_ef, rows(n_ef), Z.rank, Z.world, dist, group, ctx)
    R_cd = _stack_r(mine_cd, n_cd, rows(n_cd), Z.rank, Z.world, dist, group, ctx)
    return (_view(R_ef, (R_ef.dims[0], d["e"], d["f"])),
            _view(R_cd, (R_cd.dims[0], d["c"], d["d"])))


def _gram_sqrt(G: DeviceTensor, n: int) -> DeviceTensor:
    """R' = Lambda_+^{1/2} W^T [r; pair] from the eigendecomposition G = W Lambda W^T of a Gram
    matrix (negative rounding-noise eigenvalues clipped): R'^T R' = G to eps ||G||.
    (`rfactor="gram_eigh"`: a full n x n eigendecomposition, kept for comparison.)"""
    from .tensor import eigh_trunc

    lam, W, _ = eigh_trunc(_view(G, (n, n)), n)          # sorted by |lambda|, W: n x n
    root = np.sqrt(np.maximum(lam.to_numpy(), 0.0))      # n numbers through the host
    R = _view(W, (n, n)).permute((1, 0))
    return _scale_leg(R, 0, DeviceTensor.from_numpy(root, 1, G.ctx))


def _gram_chol(G: DeviceTensor, n: int, min_rank: int) -> DeviceTensor:
    """L [pair; r] with L L^T = G to eps ||G||: diagonally pivoted Cholesky of the Gram matrix
    on the device (`tnr_psd_factor`, csrc/pchol.cu), stopped at the numerical rank r.  L^T is an
    R factor of the matricization up to a left orthogonal gauge -- all that atrg3d.jl:58-66
    use of it.  Columns beyond the rank are zero; at least `min_rank` columns are returned so
    that the truncated bond keeps the dimension the reference gives it."""
    import ctypes as C

    L = DeviceTensor.empty((n, n), 1, G.ctx)
    r = C.c_int64(0)
    G.ctx.call("tnr_psd_factor", G.ptr, n, L.ptr, C.byref(r))
    r = min(n, max(int(r.value), int(min_rank), 1))
    return DeviceTensor(L.buf[: n * r], (n, r), None, G.ctx)


def _gram_matrix(F: TwoFactor, pair: str) -> DeviceTensor:
    """G[(p q),(p' q')] of the matricization of F = sum_i P Q with columns `pair`, from the factors."""
    up = pair.upper()
    ren = lambda lab: "".join({pair[0]: up[0], pair[1]: up[1], BOND: "I"}.get(c, c) for c in lab)  # noqa: E731
    keep_p = "".join(c for c in F.lp if c in pair or c == BOND)
    keep_q = "".join(c for c in F.lq if c in pair or c == BOND)
    PP = contract(F.P, F.lp, F.P, ren(F.lp), keep_p + ren(keep_p))
    QQ = contract(F.Q, F.lq, F.Q, ren(F.lq), keep_q + ren(keep_q))
    return contract(PP, keep_p + ren(keep_p), QQ, keep_q + ren(keep_q), pair + up)


def _projectors_gram_dealt(YD: TwoFactor, AX: TwoFactor, chi: int, shard, stats: dict):
    """rfactor="gram" on a sharded run: the two projector pairs of atrg3d.jl:58-66 are independent
    ((e f): Proj_1 / Proj_2, (c d): Proj_3 / Proj_4), so rank 0 builds one pair and rank 1 the
    other -- Gram matrices, Cholesky factors, truncated SVD -- and the chi^2 x chi projectors are
    broadcast; nothing of size chi^4 travels."""
    dist, group, rank, world = shard
    out, ranks = [], {}
    for gi, pair in enumerate(("ef", "cd")):
        owner = gi % world
        d0, d1 = YD.dim(pair[0]), YD.dim(pair[1])
        n = d0 * d1
        kk = min(chi, n)
        if rank == owner:
            Ll = _gram_chol(_gram_matrix(YD, pair), n, kk)
            Lr = _gram_chol(_gram_matrix(AX, pair), n, kk)
            ranks[pair] = [Ll.dims[1], Lr.dims[1]]
            Pa, Pb = _projectors(_view(Ll, (d0, d1, Ll.dims[1])), _view(Lr, (d0, d1, Lr.dims[1])),
                                 chi, True)
            assert Pa.dims == (d0, d1, kk) and Pb.dims == (kk, d0, d1), (Pa.dims, Pb.dims)
        else:
            Pa = DeviceTensor.empty((d0, d1, kk), None, YD.ctx)
            Pb = DeviceTensor.empty((kk, d0, d1), None, YD.ctx)
        src = dist.get_global_rank(group, owner) if group is not None else owner
        dist.broadcast(Pa.buf, src=src, group=group)
        dist.broadcast(Pb.buf, src=src, group=group)
        out += [Pa, Pb]
    stats["gram_ranks"] = ranks
    stats["projector_owners"] = [0, 1 % world]
    return out


def _r_factors_gram(F: TwoFactor, chi: int, eigh: bool = False):
    """(R_ef, R_cd) as in `_r_factors` (`eigh`: [r, pair]; default: the transposed Cholesky
    form [pair, r]), from the Gram matrices of the two matricizations, which
    follow from the factors without ever forming Z = sum_i P Q:
        G[(e f),(e' f')] = sum_{i i'} PP[.. i .. i'] QQ[.. i .. i']   (O(chi^6) flop),
    each P/Q self-contraction running over that factor's open legs that are ROWS.  R' is the
    PSD square root of G in its eigenbasis.  The projectors of atrg3d.jl:58-66 depend on R1, R2
    only through R1^T R1 and R2 R2^T (their left orthogonal gauge cancels), so this is the same
    map evaluated on Gram matrices that are exact to eps ||G||; what is lost against the TSQR
    path is accuracy in directions with sigma < 1e-8 sigma_1, and the top-chi part of R1 R2 sees
    that only through the sensitivity (s_1 / s_chi)^2 eps -- measured in the tests."""
    d = {c: F.dim(c) for c in LEGS}
    out = []
    for pair in ("ef", "cd"):
        G = _gram_matrix(F, pair)
        n = d[pair[0]] * d[pair[1]]
        if eigh:
            R = _gram_sqrt(G, n)
            out.append(_view(R, (n, d[pair[0]], d[pair[1]])))
        else:
            L = _gram_chol(G, n, min(chi, n))
            out.append(_view(L, (d[pair[0]], d[pair[1]], L.dims[1])))
    return out[0], out[1]


def _projectors(Rl: DeviceTensor, Rrt: DeviceTensor, chi: int, bond_last: bool = False):
    """atrg3d.jl:58-66 with Rl = R1 [r; p q] and Rrt = R2^T [r'; p q] (`bond_last`: [p q; r] and
    [p q; r'], the layout of the Cholesky factors):
    temp = Rl Rr,  U S V = svd_trunc(temp),  Pa[p q; k] = Rr V' S^-1/2,  Pb[k; p q] = S^-1/2 U' Rl."""
    ll, lr = ("pqr", "pqs") if bond_last else ("rpq", "spq")
    t = contract(Rl, ll, Rrt, lr, "rs")
    U, S, Vt, _ = svd_trunc(t, 1, chi)
    inv = _vec_map(S, 2, -0.5)
    Pa = _scale_leg(contract(Rrt, lr, Vt, "ks", "pqk"), 2, inv)
    Pb = _scale_leg(contract(U, "rk", Rl, ll, "kpq"), 0, inv)
    return Pa, Pb


def _squeeze(Z: _ChunkedPair, Pcd: DeviceTensor, lcd: str, Pef: DeviceTensor, lef: str, dist,
             group) -> DeviceTensor:
    """out[a b C D] = Z[a b c d e f] Pcd[..] Pef[..] (H and G of atrg3d.jl:68-69), produced chunk
    by chunk with the chunked bond as the LAST leg of `out`, all-gathered along it."""
    ctx = Z.F.ctx
    other = "b" if Z.label == "a" else "a"
    lo_ = other + "CD" + Z.label
    dC = Pcd.dims[lcd.index("C")]
    dD = Pef.dims[lef.index("D")]
    od = (Z.F.dim(other), dC, dD, Z.n)
    import torch

    out = DeviceTensor(torch.empty(max(1, math.prod(od)), dtype=torch.float64,
                                   device=ctx.torch_device), od, None, ctx)
    for lo, hi in Z.plans[Z.rank]:
        D = Z.dense(lo, hi)
        t = contract(D, "abcdef", Pcd, lcd, "abefC")
        del D
        contract(t, "abefC", Pef, lef, lo_, out=_slice_last(out, lo, hi))
    if Z.world > 1:
        from .schemes import allgather_last_leg

        allgather_last_leg(out.buf, od, group)
    return out, lo_


# ---------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------
def atrg3d_substep_factored(T: TwoFactor, chi: int, max_chunk_elems: int = 1 << 28,
                            shard: bool = False, group=None, tol: float = 1e-13,
                            block: int | None = None, rfactor: str = "tsqr") -> TwoFactor:
    """`_step!(::ATRG_3D)` (atrg3d.jl:34-83) on a TwoFactor; returns the new tensor as a
    TwoFactor with legs in the reference's order [D U; N E S W]."""
    dist, rank, world = _dist(shard, group)
    stats = {"svd": []}
    ph = _Phases(T.ctx)
    F = T.relabel()
    # U, S, V = svd_trunc(permute(T, ((2,5,6),(3,4,1))))                         atrg3d.jl:35
    st = {}
    sh = (dist, group, rank, world) if world > 1 else None
    fU, fS, fV = svd_topk_factored(F, "bef", "cda", chi, tol=tol, stats=st, block=block, shard=sh)   # [i2 i5 i6 k], [i3 i4 i1 k]
    stats["svd"].append(st)
    ph.mark("svd_T")
    US = _scale_leg(fU.clone(), 3, fS)        # C = U*S
    SV = _scale_leg(fV.clone(), 3, fS)        # B = S*V
    # M[-1 -2;-3 -4 -5 -6] := B[1 -2;-3 -4] C[-1 1;-5 -6]; as permute(M, ((2,5,6),(3,4,1))) its
    # legs are [kB i5 i6 | i3 i4 kC] = "bef|cda" with US = [i e f a], SV = [c d i b]   :41-43
    M = TwoFactor(US, "iefa", SV, "cdib", "befcda")
    st = {}
    gU, gS, gV = svd_topk_factored(M, "bef", "cda", chi, tol=tol, stats=st, block=block, shard=sh)   # [m2 m5 m6 k], [m3 m4 m1 k]
    stats["svd"].append(st)
    ph.mark("svd_M")
    del M, US, SV
    _scale_leg(gU, 3, gS, 1)                  # X = U*sqrt(S)
    _scale_leg(gV, 3, gS, 1)                  # Y = sqrt(S)*V
    # AX[-1 -2;-3 -4 -5 -6] := A[1 -2;-3 -5] X[-1 1;-4 -6]: X = gU [i d f a], A = fU [b c e i]
    # YD[-1 -2;-3 -4 -5 -6] := Y[1 -2;-3 -5] D[-1 1;-4 -6]: D = fV [d f a i], Y = gV [c e i b]
    AX = TwoFactor(gU, "idfa", fU, "bcei", LEGS)
    YD = TwoFactor(fV, "dfai", gV, "ceib", LEGS)
    other = math.prod(AX.dims) // max(1, AX.dim("a"))
    width_a = max(1, int(max_chunk_elems) // max(1, other))
    other = math.prod(YD.dims) // max(1, YD.dim("b"))
    width_b = max(1, int(max_chunk_elems) // max(1, other))
    AXc = _ChunkedPair(AX, "a", width_a, rank, world)
    YDc = _ChunkedPair(YD, "b", width_b, rank, world)
    stats["chunks"] = {"AX": [len(p) for p in AXc.plans], "YD": [len(p) for p in YDc.plans],
                       "width": (width_a, width_b), "world": world}
    dealt = rfactor == "gram" and world > 1
    if dealt:
        P1, P2, P3, P4 = _projectors_gram_dealt(YD, AX, chi, (dist, group, rank, world), stats)
    elif rfactor in ("gram", "gram_eigh"):
        R1, R3 = _r_factors_gram(YD, chi, eigh=rfactor == "gram_eigh")
        R2t, R4t = _r_factors_gram(AX, chi, eigh=rfactor == "gram_eigh")
    elif rfactor == "tsqr":
        R1, R3 = _r_factors(YDc, dist, group)      # left_orth(YD ...)   [r; 5 6], [r; 3 4]
        R2t, R4t = _r_factors(AXc, dist, group)    # right_orth(AX ...)^T
    else:
        raise ValueError(f"rfactor must be 'tsqr', 'gram' or 'gram_eigh', not {rfactor!r}")
    stats["rfactor"] = rfactor
    if not dealt:
        if rfactor == "gram":
            stats["gram_ranks"] = [R1.dims[-1], R2t.dims[-1], R3.dims[-1], R4t.dims[-1]]
        ph.mark("r_factors")
        P1, P2 = _projectors(R1, R2t, chi, rfactor == "gram")     # Proj_1 [5 6; k], Proj_2 [k; 5 6]
        P3, P4 = _projectors(R3, R4t, chi, rfactor == "gram")     # Proj_3 [3 4; k], Proj_4 [k; 3 4]
        del R1, R2t, R3, R4t
    ph.mark("projectors")
    # H[-1 -2;-3 -4] := YD[-1 -2;1 2 3 4] Proj_3[1 2;-3] Proj_1[3 4;-4]           :68
    H, lh = _squeeze(YDc, P3, "cdC", P1, "efD", dist, group)     # [a C D b]
    # G[-1 -2;-3 -4] := AX[-1 -2;1 2 3 4] Proj_4[-3;1 2] Proj_2[-4;3 4]           :69
    G, lg = _squeeze(AXc, P4, "Ccd", P2, "Def", dist, group)     # [b C D a]
    ph.mark("squeeze_H_G")
    if ph.out:
        stats["phase_s"] = ph.out
    # T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] H[-1 1;-3 -4]                        :71
    ren_h = {"a": "a", "b": BOND, "C": "c", "D": "d"}
    ren_g = {"a": BOND, "b": "b", "C": "e", "D": "f"}
    LAST_STATS.clear()
    LAST_STATS.update(stats)
    return TwoFactor(H, "".join(ren_h[c] for c in lh), G, "".join(ren_g[c] for c in lg), LEGS)


def atrg3d_step_factored(T: TwoFactor, chi: int, **kw) -> TwoFactor:
    """step!(::ATRG_3D) -- atrg3d.jl:85-97: three `_step!`s, each followed by
    permute(T, ((4,6),(2,5,1,3))), which on a TwoFactor only renames legs."""
    for _ in range(3):
        T = atrg3d_substep_factored(T, chi, **kw).permute((3, 5, 1, 4, 0, 2))
    return T
