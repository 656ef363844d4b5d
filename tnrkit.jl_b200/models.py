"""Model constructors -- host side, unchanged in spirit from /root/reference/src/models
(`src/models stays unchanged; model tensors are built on the host and uploaded once`).

Each constructor returns a host `numpy` array indexed [leg1, leg2, ...] in the reference's
leg convention (2D: V1 (x) V2 <- V3 (x) V4, 3D: D U' <- N E S' W').  `Z2Irrep` / `ZNIrrep` /
`U1Irrep` variants return a `ChargedArray`: the tensor in the charge basis (every leg graded by
the irrep label) together with the charges and arrows of its legs, from which the schemes build
the block-sparse device tensor (`symmetric.SymTensor`).  Non-abelian sectors (`DNIrrep`,
`CU1Irrep`), product sectors and the fermionic models are not built."""
from __future__ import annotations

import math

import numpy as np


class ChargedArray(np.ndarray):
    """Host tensor in the charge basis of a Z_N symmetry (or U(1): N = 0, no modulus):
    `charges[leg][i]` is the irrep label of index i of that leg, `signs[leg]` = +1 (codomain) /
    -1 (domain), negated for a dual space, and entries are nonzero only where
    sum_leg sign*charge = 0 mod N -- the information a TensorKit
    `TensorMap{Float64, Vect[ZNIrrep{N}]}` (`Vect[U1Irrep]`) carries in its spaces."""

    @classmethod
    def wrap(cls, arr, N, charges, signs):
        obj = np.asarray(arr, dtype=np.float64).view(cls)
        obj.N = int(N)
        obj.charges = [tuple(c) for c in charges]
        obj.signs = tuple(int(s) for s in signs)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.N = getattr(obj, "N", None)
        self.charges = getattr(obj, "charges", None)
        self.signs = getattr(obj, "signs", None)


class Trivial:
    """No symmetry (TensorKit.Trivial)."""


class Z2Irrep:
    """Z2 symmetry labels (TensorKitSectors.Z2Irrep)."""


class U1Irrep:
    """U(1) symmetry labels (TensorKitSectors.U1Irrep).  Charges are stored as integers; models
    with half-integer labels (six-vertex: +-1/2) store twice the charge."""


class ZNIrrep:
    """ZN symmetry labels; use ZNIrrep[N] (TensorKitSectors.ZNIrrep{N})."""

    _cache = {}

    def __class_getitem__(cls, n):
        if n not in cls._cache:
            cls._cache[n] = type(f"Z{n}Irrep", (cls,), {"N": int(n)})
        return cls._cache[n]


# src/models/ising.jl:1-8
ising_βc = math.log(1.0 + math.sqrt(2.0)) / 2.0
ising_bc = ising_βc
f_onsager = -2.10965114460820745966777928351108478082549327543540531781696107967700291143188
# src/models/ising.jl:3-7
ising_cft_exact = [1 / 8, 1, 9 / 8, 9 / 8, 2, 2, 2, 2, 17 / 8, 17 / 8, 17 / 8, 3, 3, 3, 3, 3,
                   25 / 8, 25 / 8, 25 / 8, 25 / 8, 25 / 8, 25 / 8]
ising_βc_3D = 1.0 / 4.51152469
ising_bc_3D = ising_βc_3D


def potts_βc(q):
    """src/models/potts.jl:40"""
    return math.log(1.0 + math.sqrt(q))


potts_bc = potts_βc


def _split(args, default_sym):
    sym = default_sym
    rest = list(args)
    if rest and isinstance(rest[0], type):
        sym = rest.pop(0)
    return sym, rest


def classical_ising(*args, h=0.0):
    """classical_ising([symmetry], [beta]; h)  -- src/models/ising.jl:35-66.
    Default symmetry Z2Irrep, default beta = ising_βc."""
    sym, rest = _split(args, Z2Irrep)
    beta = float(rest[0]) if rest else ising_βc
    if sym is Trivial:
        init = np.zeros((2, 2, 2, 2))
        for idx in np.ndindex(2, 2, 2, 2):
            init[idx] = math.cosh(h * beta) if sum(idx) % 2 == 0 else math.sinh(h * beta)
        b = np.diag([math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))])
        return 2.0 * np.einsum("abcd,ia,jb,ck,dl->ijkl", init, b, b, b, b)
    if sym is Z2Irrep:
        if h != 0.0:
            raise AssertionError("External magnetic field is not compatible with Z2 symmetry")
        x, y = math.cosh(beta), math.sinh(beta)
        t = np.zeros((2, 2, 2, 2))
        # coupled sector 0: uncoupled (0,0),(1,1); sector 1: (1,0),(0,1)
        b0 = [[2 * x * x, 2 * x * y], [2 * x * y, 2 * y * y]]
        b1 = [[2 * x * y, 2 * x * y], [2 * x * y, 2 * x * y]]
        for blk, pairs in ((b0, [(0, 0), (1, 1)]), (b1, [(1, 0), (0, 1)])):
            for r, (i, j) in enumerate(pairs):
                for c, (k, l) in enumerate(pairs):
                    t[i, j, k, l] = blk[r][c]
        return ChargedArray.wrap(t, 2, [(0, 1)] * 4, (1, 1, -1, -1))
    raise TypeError(f"classical_ising: unsupported symmetry {sym}")


def classical_ising_3D(*args, J=1.0):
    """classical_ising_3D([symmetry], [beta]; J)  -- src/models/ising.jl:125-165."""
    sym, rest = _split(args, Z2Irrep)
    beta = float(rest[0]) if rest else ising_βc_3D
    K = beta * J
    if sym is Trivial:
        t = np.array([[math.exp(K), math.exp(-K)], [math.exp(-K), math.exp(K)]])
        w, v = np.linalg.eigh(t)
        q = v @ np.diag(np.sqrt(w)) @ v
        O = np.zeros((2,) * 6)
        O[(0,) * 6] = 1.0
        O[(1,) * 6] = 1.0
        return np.einsum("abcdef,ia,jb,kc,ld,me,nf->ijklmn", O, q, q, q, q, q, q)
    if sym is Z2Irrep:
        x, y = math.cosh(K), math.sinh(K)
        W = np.array([[math.sqrt(x), math.sqrt(y)], [math.sqrt(x), -math.sqrt(y)]])
        t = np.einsum("ai,aj,ak,al,am,an->ijklmn", W, W, W, W, W, W)
        t = np.ascontiguousarray(np.transpose(t, (0, 3, 4, 5, 1, 2)))
        # spaces after permute(t, ((1,4),(5,6,2,3))): S (x) S' <- S (x) S (x) S' (x) S'  (ising.jl:164);
        # arrow of a leg = (+1 codomain / -1 domain) * (-1 if the space is dual)
        return ChargedArray.wrap(t, 2, [(0, 1)] * 6, (1, -1, -1, -1, 1, 1))
    raise TypeError(f"classical_ising_3D: unsupported symmetry {sym}")


def potts_tensor(q, beta):
    """src/models/potts.jl:21-30"""
    A = np.zeros((q,) * 4)
    for i, j, k, l in np.ndindex(q, q, q, q):
        E = -(int(i == j) + int(j == l) + int(k == l) + int(k == i))
        A[i, j, k, l] = math.exp(-beta * E)
    return A


def classical_potts(*args):
    """classical_potts([symmetry], q, [beta])  -- src/models/potts.jl:63-82."""
    sym, rest = _split(args, None)
    q = int(rest[0])
    beta = float(rest[1]) if len(rest) > 1 else potts_βc(q)
    if sym is None:
        sym = ZNIrrep[q]
    A = potts_tensor(q, beta)
    if sym is Trivial:
        return A
    if isinstance(sym, type) and issubclass(sym, ZNIrrep):
        if sym.N != q:
            raise AssertionError("number of irreps must match the number of states")
        w = np.exp(2j * math.pi / q)
        Wm = np.array([[w ** (r * c) for c in range(q)] for r in range(q)]) / math.sqrt(q)
        # ((P' (x) P') * A * (P (x) P)).data reshaped (q,q,q,q), column major as in Julia
        Pk = np.kron(Wm, Wm)  # row index (i,j) with j fastest <-> column-major pair (j,i)
        Am = A.reshape(q * q, q * q, order="F")
        Pcm = np.kron(Wm, Wm)
        U = Pcm.conj().T @ Am @ Pcm
        Ud = U.reshape((q, q, q, q), order="F")
        return ChargedArray.wrap(np.ascontiguousarray(Ud.real), q, [tuple(range(q))] * 4,
                                 (1, 1, -1, -1))
    raise TypeError(f"classical_potts: unsupported symmetry {sym}")


def clock_tensor(q, beta):
    """src/models/clock.jl:1-12"""
    A = np.zeros((q,) * 4)
    clock = lambda i, j: -math.cos(2.0 * math.pi / q * (i - j))  # noqa: E731
    for i, j, k, l in np.ndindex(q, q, q, q):
        E = clock(i, j) + clock(j, l) + clock(l, k) + clock(k, i)
        A[i, j, k, l] = math.exp(-beta * E)
    return A


def classical_clock(*args):
    """classical_clock([symmetry], q, beta)  -- src/models/clock.jl:26-50.  `Trivial` and
    `ZNIrrep[q]`; the reference's default, the non-abelian `DNIrrep{q}`, is outside the hot path
    (no fusion-tree recoupling on the device), so the default here is `ZNIrrep[q]`."""
    sym, rest = _split(args, None)
    q, beta = int(rest[0]), float(rest[1])
    if sym is None:
        sym = ZNIrrep[q]
    A = clock_tensor(q, beta)
    if sym is Trivial:
        return A
    if isinstance(sym, type) and issubclass(sym, ZNIrrep):
        if sym.N != q:
            raise AssertionError("number of irreps must match the number of states")
        U = np.exp(2j * math.pi / q * np.outer(np.arange(q), np.arange(q))) / math.sqrt(q)
        # Anew[-1 -2;-3 -4] := A[1 2;3 4] U[4;-4] conj(U[1;-1]) U[3;-3] conj(U[2;-2])   (clock.jl:44)
        Anew = np.einsum("ijkl,ia,jb,kc,ld->abcd", A, U.conj(), U.conj(), U, U)
        return ChargedArray.wrap(np.ascontiguousarray(Anew.real), q, [tuple(range(q))] * 4,
                                 (1, 1, -1, -1))
    raise TypeError(f"classical_clock: unsupported symmetry {sym}")


def sixvertex(*args, a=1.0, b=1.0, c=1.0):
    """sixvertex([symmetry]; a, b, c)  -- src/models/sixvertex.jl:28-48.  `Trivial` and `U1Irrep`
    (the reference's default `CU1Irrep` is non-abelian: outside the hot path; default here U1)."""
    sym, _ = _split(args, U1Irrep)
    if sym is Trivial:
        d = np.array([[a, 0, 0, 0], [0, c, b, 0], [0, b, c, 0], [0, 0, 0, a]], dtype=float)
        # TensorMap(d, C^2 (x) C^2, C^2 (x) C^2): row = i + 2 j, column = k + 2 l (column major)
        return np.ascontiguousarray(d.reshape((2, 2, 2, 2), order="F"))
    if sym is U1Irrep:
        # pspace = U1Space(-1/2 => 1, 1/2 => 1): index 0 <-> charge -1/2, index 1 <-> +1/2;
        # block(0) = [b c; c b] on {(-,+), (+,-)}, block(+-1) = [a]
        t = np.zeros((2, 2, 2, 2))
        t[0, 0, 0, 0] = t[1, 1, 1, 1] = a
        t[0, 1, 0, 1] = t[1, 0, 1, 0] = b
        t[0, 1, 1, 0] = t[1, 0, 0, 1] = c
        return ChargedArray.wrap(t, 0, [(-1, 1)] * 4, (1, 1, -1, -1))   # charges doubled
    raise TypeError(f"sixvertex: unsupported symmetry {sym}")


def _f_real(p1, p2, mu0, lam, h=0.0):
    """src/models/phi4_real.jl:1-8"""
    return math.exp(-0.5 * (p1 - p2) ** 2 - mu0 / 8.0 * (p1 ** 2 + p2 ** 2)
                    - lam / 16.0 * (p1 ** 4 + p2 ** 4) + h / 4.0 * (p1 + p2))


def phi4_real(*args):
    """phi4_real([symmetry], K, mu0, lam, [h])  -- src/models/phi4_real.jl:75-139.
    Trivial: Gauss-Hermite quadrature with K points; Z2Irrep (default): Taylor expansion of the
    hopping term, K/2 even + K/2 odd states per leg (h must be 0)."""
    sym, rest = _split(args, Z2Irrep)
    K, mu0, lam = int(rest[0]), float(rest[1]), float(rest[2])
    h = float(rest[3]) if len(rest) > 3 else 0.0
    if sym is Trivial:
        ys, ws = np.polynomial.hermite.hermgauss(K)
        f = np.array([[_f_real(ys[i], ys[j], mu0, lam, h) for j in range(K)] for i in range(K)])
        U, S, V = np.linalg.svd(f)
        rs = np.sqrt(S)
        w = ws * np.exp(ys ** 2)
        # T[i j k l] = sum_p sqrt(S_i S_j S_k S_l) w_p U[p,i] U[p,j] V[k,p] V[l,p]
        return np.einsum("p,pi,pj,kp,lp->ijkl", w, U * rs, U * rs, rs[:, None] * V,
                         rs[:, None] * V, optimize=True)
    if sym is Z2Irrep:
        if h != 0.0:
            raise AssertionError("External magnetic field is not compatible with Z2 symmetry")
        if K % 2 != 0:
            raise ValueError("K must be even to split into even/odd groups")
        from scipy.integrate import quad

        a_, b_ = (4.0 + mu0) / 2.0, lam / 4.0
        moments = np.zeros(4 * (K - 1) + 1)
        for n in range(0, 4 * (K - 1) + 1, 2):
            moments[n] = quad(lambda x, n=n: math.exp(-a_ * x * x - b_ * x ** 4) * x ** n,
                              -np.inf, np.inf, epsabs=0.0, epsrel=1e-12, limit=400)[0]
        logfact = np.array([math.lgamma(s + 1.0) for s in range(K)])
        t = np.zeros((K,) * 4)
        for s in np.ndindex(K, K, K, K):
            n = sum(s)
            if n % 2:
                continue
            t[s] = moments[n] / math.exp(0.5 * sum(logfact[x] for x in s))
        perm = list(range(0, K, 2)) + list(range(1, K, 2))     # evens, then odds
        t = t[np.ix_(perm, perm, perm, perm)]
        ch = (0,) * (K // 2) + (1,) * (K // 2)
        return ChargedArray.wrap(np.ascontiguousarray(t), 2, [ch] * 4, (1, 1, -1, -1))
    raise TypeError(f"phi4_real: unsupported symmetry {sym}")


def _f_complex(r1, c1, r2, c2, mu0, lam):
    """src/models/phi4_complex.jl:1-7"""
    return math.exp(-0.5 * ((r1 - r2) ** 2 + (c1 - c2) ** 2)
                    - mu0 / 8.0 * (r1 ** 2 + c1 ** 2 + r2 ** 2 + c2 ** 2)
                    - lam / 16.0 * ((r1 ** 2 + c1 ** 2) ** 2 + (r2 ** 2 + c2 ** 2) ** 2))


def phi4_complex(*args):
    """phi4_complex([symmetry], K, mu0, lam)  -- src/models/phi4_complex.jl:152-160, 219-281.
    Trivial: Gauss-Hermite quadrature in Re/Im (bond dimension K^2); U1Irrep (default): Taylor
    expansion in phi, phi^* with charges q = a - B, a, B in 0..K-1 (bond dimension K^2, 2K-1
    sectors).  The Z2 x Z2 variant (product sector) is not built."""
    sym, rest = _split(args, U1Irrep)
    K, mu0, lam = int(rest[0]), float(rest[1]), float(rest[2])
    if sym is Trivial:
        ys, ws = np.polynomial.hermite.hermgauss(K)
        N = K * K
        # fmatrix_complex: index (i-1)K + j  <->  phi = ys[i] + i ys[j]
        f = np.empty((N, N))
        for i in range(K):
            for j in range(K):
                for k in range(K):
                    for l in range(K):
                        f[i * K + j, k * K + l] = _f_complex(ys[i], ys[j], ys[k], ys[l], mu0, lam)
        w = np.array([ws[a] * ws[b] * math.exp(ys[a] ** 2 + ys[b] ** 2)
                      for a in range(K) for b in range(K)])
        U, S, V = np.linalg.svd(f)
        rs = np.sqrt(S)
        full = np.einsum("p,pi,pj,kp,lp->ijkl", w, U * rs, U * rs, rs[:, None] * V, rs[:, None] * V,
                         optimize=True)
        # phi4_complex_tensor computes the entry of the sorted index tuple and assigns it to all
        # 24 permutations (phi4_complex.jl:117-137)
        idx = np.sort(np.stack(np.meshgrid(*[np.arange(N)] * 4, indexing="ij"), axis=-1), axis=-1)
        return np.ascontiguousarray(full[idx[..., 0], idx[..., 1], idx[..., 2], idx[..., 3]])
    if sym is U1Irrep:
        if K % 2 != 0:
            raise ValueError("K must be even")
        from scipy.integrate import quad

        a_, b_ = 2.0 + mu0 / 2.0, lam / 4.0
        nmax = 8 * (K - 1) + 1
        moments = np.zeros(nmax + 1)
        for n in range(nmax + 1):
            moments[n] = quad(lambda r, n=n: math.exp(n * math.log(r) - a_ * r * r - b_ * r ** 4)
                              if r > 0.0 else (1.0 if n == 0 else 0.0),
                              0.0, np.inf, epsabs=0.0, epsrel=1e-13, limit=400)[0]
        logfact = np.array([math.lgamma(x + 1.0) for x in range(K)])
        # basis of W = fuse(V1, V2): states (a, B) with charge q = a - B, ordered by (q, a)
        states = sorted(((a - B, a, B) for a in range(K) for B in range(K)))
        ch = tuple(q for q, _, _ in states)
        N = len(states)
        pw = np.array([a + B for _, a, B in states])
        lf = np.array([logfact[a] + logfact[B] for _, a, B in states])
        q = np.array(ch)
        sp = pw[:, None, None, None] + pw[None, :, None, None] + pw[None, None, :, None] + pw[None, None, None, :]
        ld = 0.5 * (math.log(2.0) * sp + lf[:, None, None, None] + lf[None, :, None, None]
                    + lf[None, None, :, None] + lf[None, None, None, :])
        ok = (q[:, None, None, None] + q[None, :, None, None]
              == q[None, None, :, None] + q[None, None, None, :])
        t = np.where(ok, 2.0 * math.pi * moments[sp + 1] / np.exp(ld), 0.0)
        assert t.shape == (N,) * 4
        return ChargedArray.wrap(np.ascontiguousarray(t), 0, [ch] * 4, (1, 1, -1, -1))
    raise TypeError(f"phi4_complex: unsupported symmetry {sym}")


XY_βc = 1.1199   # src/models/XY.jl:15 ("This is an approximation!")
XY_bc = XY_βc


def classical_XY(*args):
    """classical_XY([U1Irrep], [beta], charge_trunc)  -- src/models/XY.jl:40-55 + algebraic_initialization
    (:1-12): legs carry the U(1) charges -charge_trunc..charge_trunc (one state each), the fusion
    tensor m[u; A B] = 1 where u = A + B, bonds weigh charge q with the Bessel function I_q(beta).
    Only `U1Irrep` (the reference's default `CU1Irrep` = O(2) is non-abelian)."""
    from scipy.special import iv

    sym, rest = _split(args, U1Irrep)
    if sym is not U1Irrep:
        raise TypeError(f"classical_XY: unsupported symmetry {sym}")
    beta = float(rest[0]) if len(rest) > 1 else XY_βc
    c = int(rest[-1])
    q = np.arange(-c, c + 1)
    n = q.size
    m = (q[:, None, None] == q[None, :, None] + q[None, None, :]).astype(float)   # m[u; A B]
    w = iv(q, beta)
    # T[l u; d r] := m[u;Au Bu] bond[Au;Ad] bond[Bu;Bd] m[Bd;Cu r] m'[l Ad;Du] bond[Du;Dd]
    #                bond[Cu;Cd] m'[Dd Cd;d]       with m'[x y; z] = m[z; x y]
    mw = m * w[None, :, None] * w[None, None, :]            # m[u; A B] I_A I_B
    md = m * w[:, None, None]                               # m[D; l A] I_D
    mc = m * w[None, :, None]                               # m[B; C r] I_C
    t = np.einsum("uAB,BCr,DlA,dDC->ludr", mw, mc, md, m, optimize=True)
    assert t.shape == (n,) * 4
    return ChargedArray.wrap(np.ascontiguousarray(t), 0, [tuple(int(x) for x in q)] * 4,
                             (1, 1, -1, -1))
