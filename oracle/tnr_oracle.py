"""CPU oracle for the TNRKit coarse-graining hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy (LAPACK via numpy.linalg) restatement of the
reference algorithms in /root/reference/src (TNRKit.jl v0.5.1).  It is the
checker for the CUDA path; it is never the thing shipped or measured.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs may import it.  The product package (tnrkit.jl_b200) never
imports it and fails loudly when libtnrcuda.so is missing.

Third-party algorithms the reference delegates to (absent from /root/reference,
pinned in its Project.toml [compat]): TensorKit 0.16.2 (`@tensor`, `permute`,
`svd_trunc`, `eigh_trunc`, `left_orth`, `right_orth`), MatrixAlgebraKit 0.6.1
(`truncrank`: keep the `howmany` largest values by abs; truncation error =
2-norm of the discarded values).  For `Trivial` (dense) tensors these reduce to
dense index permutation, pairwise contraction, LAPACK SVD / symmetric
eigendecomposition and thin QR, which is what is restated here.  For abelian
`Z2Irrep` / `ZNIrrep{N}` tensors the block structure is restated in
`oracle/sym_oracle.py`.

Parity pin: the oracle is checked (tests/test_oracle_golden.py) against every
golden number the reference holds for this path: the README quick-start value
(BTRG chi=16, 25 iterations, README.md:81), f_onsager with the rtol of each
scheme's testset in test/schemes.jl, and f_benchmark3D for the 3D schemes.

Conventions: a TensorMap with codomain legs (1..N1) and domain legs
(N1+1..N1+N2) is a numpy array whose axis k-1 is reference leg k.  numpy
`transpose(T, axes)` has the same meaning as TensorKit `permute(T, p)` with
p = axes+1 (new leg k is old leg p[k]).
"""
from __future__ import annotations

import math
import numpy as np

# einsum path: greedy pairwise contraction WITHOUT numpy's default cap on intermediate size (the cap
# makes a 4-operand network with input legs larger than chi fall back to one naive loop)
_OPT = ("greedy", 1 << 40)

# ----------------------------------------------------------------------------
# constants  (src/models/ising.jl:1-8, src/models/potts.jl:33)
# ----------------------------------------------------------------------------
ising_bc = math.log(1.0 + math.sqrt(2.0)) / 2.0
f_onsager = -2.10965114460820745966777928351108478082549327543540531781696107967700291143188
ising_bc_3D = 1.0 / 4.51152469
f_benchmark3D = -3.507  # test/schemes.jl:11


def potts_bc(q):
    return math.log(1.0 + math.sqrt(q))


# ----------------------------------------------------------------------------
# models (host-side; src/models stays unchanged in the reference)
# ----------------------------------------------------------------------------
def classical_ising(beta=ising_bc, h=0.0):
    """src/models/ising.jl:42-53 (Trivial)."""
    init = np.zeros((2, 2, 2, 2))
    for idx in np.ndindex(2, 2, 2, 2):
        # Julia indices are 1-based: mod(i+j+k+l,2) has the same parity 0-based
        init[idx] = math.cosh(h * beta) if sum(idx) % 2 == 0 else math.sinh(h * beta)
    b = np.diag([math.sqrt(math.cosh(beta)), math.sqrt(math.sinh(beta))])
    return 2.0 * np.einsum("abcd,ia,jb,ck,dl->ijkl", init, b, b, b, b)


def classical_ising_z2basis(beta=ising_bc):
    """src/models/ising.jl:54-66 (Z2Irrep) written out as a dense array in the
    charge basis (index 0 = even, 1 = odd).  Blocks: coupled sector 0 holds
    rows/cols {(0,0),(1,1)}, sector 1 holds {(0,1),(1,0)} (TensorKit fusion-tree
    order: uncoupled sectors iterate first leg fastest)."""
    x, y = math.cosh(beta), math.sinh(beta)
    t = np.zeros((2, 2, 2, 2))
    even = [(0, 0), (1, 1)]
    odd = [(1, 0), (0, 1)]
    b0 = np.array([[2 * x * x, 2 * x * y], [2 * x * y, 2 * y * y]])
    b1 = np.array([[2 * x * y, 2 * x * y], [2 * x * y, 2 * x * y]])
    for r, (i, j) in enumerate(even):
        for c, (k, l) in enumerate(even):
            t[i, j, k, l] = b0[r, c]
    for r, (i, j) in enumerate(odd):
        for c, (k, l) in enumerate(odd):
            t[i, j, k, l] = b1[r, c]
    return t


def classical_ising_3D(beta=ising_bc_3D, J=1.0):
    """src/models/ising.jl:129-146 (Trivial)."""
    K = beta * J
    t = np.array([[math.exp(K), math.exp(-K)], [math.exp(-K), math.exp(K)]])
    w, v = np.linalg.eigh(t)
    # NB the reference writes r.vectors * sqrt(D) * r.vectors (no transpose);
    # for this 2x2 symmetric t the eigenvector matrix is symmetric up to column
    # signs, restated literally here.
    q = v @ np.diag(np.sqrt(w)) @ v
    O = np.zeros((2,) * 6)
    O[(0,) * 6] = 1.0
    O[(1,) * 6] = 1.0
    return np.einsum("abcdef,ia,jb,kc,ld,me,nf->ijklmn", O, q, q, q, q, q, q)


def classical_ising_3D_z2basis(beta=ising_bc_3D, J=1.0):
    """src/models/ising.jl:147-165 (Z2Irrep), dense array in the charge basis."""
    x, y = math.cosh(beta * J), math.sinh(beta * J)
    W = np.array([[math.sqrt(x), math.sqrt(y)], [math.sqrt(x), -math.sqrt(y)]])
    t = np.einsum("ai,aj,ak,al,am,an->ijklmn", W, W, W, W, W, W)
    # permute(t, ((1,4),(5,6,2,3)))
    return np.transpose(t, (0, 3, 4, 5, 1, 2))


def classical_potts(q=3, beta=None):
    """src/models/potts.jl:21-30 (Trivial)."""
    beta = potts_bc(q) if beta is None else beta
    A = np.zeros((q,) * 4)
    for i, j, k, l in np.ndindex(q, q, q, q):
        E = -(int(i == j) + int(j == l) + int(k == l) + int(k == i))
        A[i, j, k, l] = math.exp(-beta * E)
    return A


# ----------------------------------------------------------------------------
# truncated factorizations  (TensorKit svd_trunc / eigh_trunc with truncrank)
# ----------------------------------------------------------------------------
def svd_trunc(T, ncod, chi):
    """svd_trunc(T; trunc=truncrank(chi)) of a TensorMap with `ncod` codomain legs.
    Returns U (cod.., k), s (k,), Vh (k, dom..), eps."""
    cod, dom = T.shape[:ncod], T.shape[ncod:]
    M = T.reshape(int(np.prod(cod)), int(np.prod(dom)))
    U, s, Vh = np.linalg.svd(M, full_matrices=False)
    k = min(chi, s.shape[0])
    eps = float(np.linalg.norm(s[k:]))
    return U[:, :k].reshape(*cod, k), s[:k].copy(), Vh[:k, :].reshape(k, *dom), eps


def eigh_trunc(MM, ncod, chi):
    """eigh_trunc!(MM; trunc=truncrank(chi)): keep the chi eigenvalues largest by
    abs; eps = 2-norm of the discarded eigenvalues."""
    cod = MM.shape[:ncod]
    n = int(np.prod(cod))
    M = MM.reshape(n, n)
    w, V = np.linalg.eigh(M)
    order = np.argsort(-np.abs(w), kind="stable")
    k = min(chi, n)
    keep, drop = order[:k], order[k:]
    eps = float(np.linalg.norm(w[drop]))
    return w[keep], V[:, keep].reshape(*cod, k), eps


def project_hermitian(MM, ncod):
    cod = MM.shape[:ncod]
    n = int(np.prod(cod))
    M = MM.reshape(n, n)
    return (0.5 * (M + M.T)).reshape(MM.shape)


def pseudopow(s, a, tol=np.finfo(float).eps ** 0.75):
    """src/schemes/btrg.jl:51-60."""
    s = np.asarray(s, dtype=float)
    out = s.copy()
    m = ~(s < tol)
    out[m] = s[m] ** a
    return out


# ----------------------------------------------------------------------------
# schemes
# ----------------------------------------------------------------------------
class Scheme:
    pass


class TRG(Scheme):
    """src/schemes/trg.jl."""

    def __init__(self, T):
        self.T = np.array(T, dtype=float)

    def step(self, chi):
        T = self.T
        # SVD12 (src/utility/projectors.jl:213-219): U*sqrt(s), sqrt(s)*V
        U, s, V, _ = svd_trunc(T, 2, chi)
        A, B = U * np.sqrt(s), np.sqrt(s)[:, None, None] * V
        Tp = np.transpose(T, (1, 3, 0, 2))  # transpose(T, ((2,4),(1,3)))
        U, s, V, _ = svd_trunc(Tp, 2, chi)
        C, D = U * np.sqrt(s), np.sqrt(s)[:, None, None] * V
        # T[-1 -2; -3 -4] := D[-2; 1 2] * B[-1; 4 1] * C[4 3; -3] * A[3 2; -4]
        self.T = np.einsum("bpq,asp,src,rqd->abcd", D, B, C, A, optimize=_OPT)

    def finalize(self):
        n = abs(np.einsum("abba->", self.T))  # T[1 2; 2 1]
        self.T = self.T / n
        return n


class BTRG(Scheme):
    """src/schemes/btrg.jl."""

    def __init__(self, T, k=-0.5):
        self.T = np.array(T, dtype=float)
        self.S1 = np.eye(T.shape[1])
        self.S2 = np.eye(T.shape[0])
        self.k = k

    def step(self, chi):
        T, k = self.T, self.k
        U, S, V, _ = svd_trunc(T, 2, chi)
        Sa, Sb = pseudopow(S, (1 - k) / 2), pseudopow(S, k)
        A, B, S1n = U * Sa, Sa[:, None, None] * V, np.diag(Sb)
        U, S, V, _ = svd_trunc(np.transpose(T, (2, 0, 3, 1)), 2, chi)  # ((3,1),(4,2))
        Sa, Sb = pseudopow(S, (1 - k) / 2), pseudopow(S, k)
        C, D, S2n = U * Sa, Sa[:, None, None] * V, np.diag(Sb)
        # T := D[-1;4 7] S1[1;7] B[-2;1 3] S2[3;2] C[8 2;-4] S1[8;5] A[6 5;-3] S2[4;6]
        self.T = np.einsum("aqt,it,bik,kj,ujd,uo,poc,qp->abcd",
                           D, self.S1, B, self.S2, C, self.S1, A, self.S2, optimize=_OPT)
        self.S1, self.S2 = S1n, S2n

    def finalize(self):
        # T[1 2; 4 3] * S1[4; 2] * S2[3; 1]
        n = abs(np.einsum("abdc,db,ca->", self.T, self.S1, self.S2))
        self.T = self.T / n
        return n


class HOTRG(Scheme):
    """src/schemes/hotrg.jl."""

    def __init__(self, T):
        self.T = np.array(T, dtype=float)

    @staticmethod
    def _xproj(A1, A2, chi):
        # hotrg.jl:102-106: MM := A2[-1 5;1 2] A1[-2 3;5 4] conj(A2[-3 6;1 2]) conj(A1[-4 3;6 4])
        X = np.tensordot(A2, A2, axes=([2, 3], [2, 3]))      # [a e c f]
        Y = np.tensordot(A1, A1, axes=([1, 3], [1, 3]))      # [b e d f]
        MM = np.tensordot(X, Y, axes=([1, 3], [1, 3]))       # [a c b d]
        MM = np.transpose(MM, (0, 2, 1, 3))
        _, U, e = eigh_trunc(project_hermitian(MM, 2), 2, chi)
        # hotrg.jl:110-114: MM := conj(A2[2 5;1 -1]) conj(A1[4 3;5 -2]) A2[2 6;1 -3] A1[4 3;6 -4]
        X = np.tensordot(A2, A2, axes=([0, 2], [0, 2]))      # [e a f c]
        Y = np.tensordot(A1, A1, axes=([0, 1], [0, 1]))      # [e b f d]
        MM = np.tensordot(X, Y, axes=([0, 2], [0, 2]))       # [a c b d]
        MM = np.transpose(MM, (0, 2, 1, 3))
        _, U2, e2 = eigh_trunc(project_hermitian(MM, 2), 2, chi)
        return (U2, e2) if e > e2 else (U, e)

    @staticmethod
    def _yproj(A1, A2, chi):
        # hotrg.jl:137-141: MM := A1[1 -1;2 5] A2[5 -2;4 3] conj(A1[1 -3;2 6]) conj(A2[6 -4;4 3])
        X = np.tensordot(A1, A1, axes=([0, 2], [0, 2]))      # [a e c f]
        Y = np.tensordot(A2, A2, axes=([2, 3], [2, 3]))      # [e b f d]
        MM = np.tensordot(X, Y, axes=([1, 3], [0, 2]))       # [a c b d]
        MM = np.transpose(MM, (0, 2, 1, 3))
        _, U, e = eigh_trunc(project_hermitian(MM, 2), 2, chi)
        # hotrg.jl:145-149: MM := conj(A1[1 2;-1 5]) conj(A2[5 4;-2 3]) A1[1 2;-3 6] A2[6 4;-4 3]
        X = np.tensordot(A1, A1, axes=([0, 1], [0, 1]))      # [a e c f]
        Y = np.tensordot(A2, A2, axes=([1, 3], [1, 3]))      # [e b f d]
        MM = np.tensordot(X, Y, axes=([1, 3], [0, 2]))       # [a c b d]
        MM = np.transpose(MM, (0, 2, 1, 3))
        _, U2, e2 = eigh_trunc(project_hermitian(MM, 2), 2, chi)
        return (U2, e2) if e > e2 else (U, e)

    @staticmethod
    def _step_y(A1, A2, Ux):
        # hotrg.jl:57-58: T[-1 -2;-3 -4] := conj(Ux[1 2;-1]) Ux[3 4;-4] A2[1 5;-3 3] A1[2 -2;5 4]
        W = np.tensordot(Ux, A2, axes=([0], [0]))            # [j a m c k]
        W = np.tensordot(W, A1, axes=([0, 2], [0, 2]))       # [a c k b l]
        W = np.tensordot(W, Ux, axes=([2, 4], [0, 1]))       # [a c b d]
        return np.transpose(W, (0, 2, 1, 3))

    @staticmethod
    def _step_x(A1, A2, Uy):
        # hotrg.jl:79-80: T := A1[-1 1;3 5] A2[5 2;4 -4] conj(Uy[1 2;-2]) Uy[3 4;-3]
        W = np.tensordot(A1, Uy, axes=([1], [0]))            # [a k m j b]
        W = np.tensordot(W, A2, axes=([2, 3], [0, 1]))       # [a k b l d]
        W = np.tensordot(W, Uy, axes=([1, 3], [0, 1]))       # [a b d c]
        return np.transpose(W, (0, 1, 3, 2))

    def step(self, chi):
        T = self.T
        Ux, _ = self._xproj(T, T, chi)
        T = self._step_y(T, T, Ux)
        Uy, _ = self._yproj(T, T, chi)
        self.T = self._step_x(T, T, Uy)

    finalize = TRG.finalize


class ATRG(Scheme):
    """src/schemes/atrg.jl."""

    def __init__(self, T):
        self.T = np.array(T, dtype=float)

    def _step(self, chi):
        T = self.T
        A, S, B, _ = svd_trunc(np.transpose(T, (0, 2, 1, 3)), 2, chi)  # ((1,3),(2,4))
        C, D = A.copy(), B.copy()
        B = S[:, None, None] * B
        C = C * S
        # M[-1 -2; -3 -4] := B[-3; 1 -4] * C[-1 1; -2]
        M = np.einsum("cid,aib->abcd", B, C)
        X, S, Y, _ = svd_trunc(np.transpose(M, (0, 2, 1, 3)), 2, chi)
        X, Y = X * np.sqrt(S), np.sqrt(S)[:, None, None] * Y
        # Q[-1 -2; -3 -4] := A[3 -3; 2] * D[1; -2 4] * X[4 2; -4] * Y[-1 1; 3]
        AXm = np.tensordot(A, X, axes=([2], [1]))             # [k c l d]
        YDm = np.tensordot(Y, D, axes=([1], [0]))             # [a k b l]
        Q = np.tensordot(YDm, AXm, axes=([1, 3], [0, 2]))     # [a b c d]
        H, S, G, _ = svd_trunc(Q, 2, chi)
        H, G = H * np.sqrt(S), np.sqrt(S)[:, None, None] * G
        # T[-1 -2; -3 -4] := G[-1; -3 1] * H[1 -2; -4]
        self.T = np.einsum("aci,ibd->abcd", G, H)

    def step(self, chi):
        self._step(chi)
        self.T = np.transpose(self.T, (1, 3, 0, 2))  # ((2,4),(1,3))
        self._step(chi)
        self.T = np.transpose(self.T, (2, 0, 3, 1))  # ((3,1),(4,2))

    finalize = TRG.finalize


class HOTRG_3D(Scheme):
    """src/schemes/hotrg3d.jl (bosonic: the twists are identities)."""

    def __init__(self, T):
        self.T = np.array(T, dtype=float)

    @staticmethod
    def _MMdag(A1, A2):
        # hotrg3d.jl:57-61.  A[z z2; Y X y x]
        m2 = np.tensordot(A2, A2, axes=([1, 2, 3, 4], [1, 2, 3, 4]))  # [z x2 z' x2']
        m1 = np.tensordot(A1, A1, axes=([0, 2, 3, 4], [0, 2, 3, 4]))  # [z x1 z' x1']
        MM = np.einsum("zbwd,zawc->abcd", m2, m1)  # [x1 x2; x1' x2']
        return project_hermitian(MM, 2)

    @staticmethod
    def _MdagM(A1, A2):
        # hotrg3d.jl:78-82: open legs are the 4th leg (x') here
        m2 = np.tensordot(A2, A2, axes=([1, 2, 4, 5], [1, 2, 4, 5]))  # [z x2 z' x2']
        m1 = np.tensordot(A1, A1, axes=([0, 2, 4, 5], [0, 2, 4, 5]))
        MM = np.einsum("zbwd,zawc->abcd", m2, m1)
        return project_hermitian(MM, 2)

    @classmethod
    def _xproj(cls, A1, A2, chi):
        _, U, e = eigh_trunc(cls._MMdag(A1, A2), 2, chi)
        _, U2, e2 = eigh_trunc(cls._MdagM(A1, A2), 2, chi)
        return (U2, e2) if e > e2 else (U, e)

    @classmethod
    def _yproj(cls, A1, A2, chi):
        perm = (0, 1, 3, 2, 5, 4)  # ((1,2),(4,3,6,5))
        return cls._xproj(np.transpose(A1, perm), np.transpose(A2, perm), chi)

    @staticmethod
    def _contract(A1, A2, Ux, Uy):
        # hotrg3d.jl:116-120
        # T[-1 -2;-3 -4 -5 -6] := conj(Ux[x1 x2;-6]) Ux[x1' x2';-4] conj(Uy[y1 y2;-5])
        #     Uy[y1' y2';-3] A1[-1 z; y1' x1' y1 x1] A2[z -2; y2' x2' y2 x2]
        Q = np.tensordot(A1, Ux, axes=([5], [0]))          # [a z y1' x1' y1 x2 f]
        P = np.tensordot(A2, Ux, axes=([3], [1]))          # [z b y2' y2 x2 x1' d]
        R = np.tensordot(Q, P, axes=([1, 3, 5], [0, 5, 4]))  # [a y1' y1 f b y2' y2 d]
        R = np.tensordot(R, Uy, axes=([2, 6], [0, 1]))     # [a y1' f b y2' d e]
        R = np.tensordot(R, Uy, axes=([1, 4], [0, 1]))     # [a f b d e c]
        return np.transpose(R, (0, 2, 5, 3, 4, 1))

    def _step(self, chi):
        T = self.T
        Ux, _ = self._xproj(T, T, chi)
        Uy, _ = self._yproj(T, T, chi)
        self.T = self._contract(T, T, Ux, Uy)

    def step(self, chi):
        for _ in range(3):
            self._step(chi)
            self.T = np.transpose(self.T, (5, 3, 1, 2, 0, 4))  # ((6,4),(2,3,1,5))

    def finalize(self):
        n = abs(np.einsum("aabcbc->", self.T))  # T[1 1; 2 3 2 3]
        self.T = self.T / n
        return n


class ATRG_3D(Scheme):
    """src/schemes/atrg3d.jl."""

    def __init__(self, T):
        self.T = np.array(T, dtype=float)

    @staticmethod
    def _p41_23(X):
        return np.transpose(X, (3, 0, 1, 2))  # permute(X, ((4,1),(2,3)))

    def _step(self, chi):
        T = self.T
        p = self._p41_23
        perm = (1, 4, 5, 2, 3, 0)  # ((2,5,6),(3,4,1))
        U, S, V, _ = svd_trunc(np.transpose(T, perm), 3, chi)
        A, D = p(U), p(V)
        C, B = p(U * S), p(S[:, None, None, None] * V)
        # M[-1 -2;-3 -4 -5 -6] := B[1 -2;-3 -4] * C[-1 1;-5 -6]
        M = np.einsum("ibcd,aief->abcdef", B, C)
        U, S, V, _ = svd_trunc(np.transpose(M, perm), 3, chi)
        rs = np.sqrt(S)
        X, Y = p(U * rs), p(rs[:, None, None, None] * V)
        AX = np.einsum("ibce,aidf->abcdef", A, X)  # A[1 -2;-3 -5] X[-1 1;-4 -6]
        YD = np.einsum("ibce,aidf->abcdef", Y, D)  # Y[1 -2;-3 -5] D[-1 1;-4 -6]
        sh = AX.shape

        def mat(Z, rows, cols):
            Zp = np.transpose(Z, rows + cols)
            return Zp.reshape(int(np.prod([sh[i] for i in rows])), -1)

        # left_orth -> R of thin QR; right_orth -> L of thin LQ
        R1 = np.linalg.qr(mat(YD, (0, 1, 2, 3), (4, 5)), mode="r")          # [r; 5 6]
        R2 = np.linalg.qr(mat(AX, (4, 5), (0, 1, 2, 3)).T, mode="r").T      # [5 6; r]
        R3 = np.linalg.qr(mat(YD, (0, 1, 4, 5), (2, 3)), mode="r")          # [r; 3 4]
        R4 = np.linalg.qr(mat(AX, (2, 3), (0, 1, 4, 5)).T, mode="r").T      # [3 4; r]

        def projectors(Rl, Rr):
            t = Rl @ Rr
            Uu, Ss, Vv = np.linalg.svd(t, full_matrices=False)
            k = min(chi, Ss.shape[0])
            Uu, Ss, Vv = Uu[:, :k], Ss[:k], Vv[:k, :]
            inv = pseudopow(Ss, -0.5)
            Pa = Rr @ Vv.T * inv              # [pair; k]
            Pb = (inv[:, None] * Uu.T) @ Rl   # [k; pair]
            return Pa, Pb

        P1, P2 = projectors(R1, R2)
        P3, P4 = projectors(R3, R4)
        d3, d4, d5, d6 = sh[2], sh[3], sh[4], sh[5]
        P1 = P1.reshape(d5, d6, -1)
        P2 = P2.reshape(-1, d5, d6)
        P3 = P3.reshape(d3, d4, -1)
        P4 = P4.reshape(-1, d3, d4)
        # H[-1 -2;-3 -4] := YD[-1 -2;1 2 3 4] Proj_3[1 2;-3] Proj_1[3 4;-4]
        H = np.tensordot(np.tensordot(YD, P3, axes=([2, 3], [0, 1])), P1, axes=([2, 3], [0, 1]))
        # G[-1 -2;-3 -4] := AX[-1 -2;1 2 3 4] Proj_4[-3;1 2] Proj_2[-4;3 4]
        G = np.tensordot(np.tensordot(AX, P4, axes=([2, 3], [1, 2])), P2, axes=([2, 3], [1, 2]))
        # T[-1 -2;-3 -4 -5 -6] := G[1 -2;-5 -6] * H[-1 1;-3 -4]
        self.T = np.einsum("ibef,aicd->abcdef", G, H)

    def step(self, chi):
        for _ in range(3):
            self._step(chi)
            self.T = np.transpose(self.T, (3, 5, 1, 4, 0, 2))  # ((4,6),(2,5,1,3))

    finalize = HOTRG_3D.finalize


# ----------------------------------------------------------------------------
# driver  (src/schemes/tnrscheme.jl:31-57, src/utility/stopping.jl,
#          src/utility/free_energy.jl)
# ----------------------------------------------------------------------------
def run(scheme, chi, maxiter, finalize_beginning=True):
    """run!(scheme, truncrank(chi), maxiter(n)) -> list of norms."""
    data = []
    if finalize_beginning:
        data.append(scheme.finalize())
    steps = 0
    crit = True
    while crit:
        scheme.step(chi)
        data.append(scheme.finalize())
        steps += 1
        crit = steps < maxiter
    return data


def free_energy(data, beta, scalefactor=2.0, initial_size=1.0):
    lnz = 0.0
    x = 1.0 - math.log(initial_size) / math.log(scalefactor)
    for i, z in enumerate(data, start=1):
        lnz += math.log(z) * scalefactor ** (x - i)
    return -lnz / beta


def hotrg3d_chunk_contract(Qk, Pk):
    """The dominant contraction of HOTRG_3D._contract for one (f, d) pair of open bonds:
    R[(a y1' y1), (y2 b y2')] = sum_(z x1' x2) Qk[(z x1' x2), (a y1' y1)] Pk[(z x1' x2), ...]
    (hotrg3d.jl:116-120 after absorbing Ux into A1 and A2).  Used by bench.py's cpu_baseline
    leg as the bounded CPU sample of the workload."""
    return Qk.T @ Pk


def finalize_two_by_two(scheme):
    """src/utility/finalize.jl:17-25 (2x2 unit-cell norm); BTRG: finalize.jl:27-42,
    T[11 1;9 8] S2[8;2] T[2 6;10 11] S1[3;6] T[7 10;3 12] S2[4;7] T[12 9;5 4] S1[5;1]."""
    T = scheme.T
    if isinstance(scheme, BTRG):
        # labels 1..12 -> a..l
        n = abs(np.einsum("kaih,hb,bfjk,cf,gjcl,dg,lied,ea->", T, scheme.S2, T, scheme.S1, T,
                          scheme.S2, T, scheme.S1, optimize=_OPT))
    else:
        n = abs(np.einsum("gaed,dbfg,cfbh,heac->", T, T, T, T, optimize=_OPT))
    f = n ** 0.25
    scheme.T = T / f
    return f


# ---- observables of the fixed-point tensor (src/utility/cft.jl) and their finalizers ----------
ising_cft_exact = [1 / 8, 1, 9 / 8, 9 / 8, 2, 2, 2, 2, 17 / 8, 17 / 8, 17 / 8, 3, 3, 3, 3, 3,
                   25 / 8, 25 / 8, 25 / 8, 25 / 8, 25 / 8, 25 / 8]   # src/models/ising.jl:3-7


def _unit_tensor(scheme):
    """BTRG: T_unit[-1 -2;-3 -4] := T[1 2;-3 -4] S1[-2;2] S2[-1;1] (cft.jl:44-45, 319-320,
    386-387); every other scheme: T itself."""
    if isinstance(scheme, BTRG):
        return np.einsum("abcd,yb,xa->xycd", scheme.T, scheme.S1, scheme.S2, optimize=_OPT)
    return scheme.T


def transfer_matrix(scheme, unitcell=1):
    """ncon(fill(T, u), [[i, -i, -(i+u), i+1] ...; last leg 4 -> 1]) permuted to
    ((1..u), (u+1..2u)) -- cft.jl:7-16 / 280-295: a ring of u tensors along legs 1/4, matrix
    from the legs 2 (rows) to the legs 3 (columns)."""
    T = _unit_tensor(scheme)
    R = T                                   # [a, b1.., c1.., d]
    for i in range(1, unitcell):
        # contract the running leg d with leg 1 of the next tensor
        R = np.tensordot(R, T, axes=([R.ndim - 1], [0]))      # [a, b.., c.., b', c', d']
        nb = i
        order = [0] + list(range(1, 1 + nb)) + [1 + 2 * nb] + list(range(1 + nb, 1 + 2 * nb)) + \
            [2 + 2 * nb, 3 + 2 * nb]
        R = np.transpose(R, order)          # [a, b.., b', c.., c', d']
    R = np.trace(R, axis1=0, axis2=R.ndim - 1)
    n = int(np.prod(R.shape[:unitcell]))
    return R.reshape(n, -1)


def cft_data(scheme, v=1, unitcell=1, is_real=True):
    """cft.jl:5-37 (TNRScheme) and 39-73 (BTRG): scaling dimensions from the eigenvalues of the
    transfer matrix."""
    data = np.linalg.eigvals(transfer_matrix(scheme, unitcell)).astype(complex)
    data = data[np.argsort(-np.abs(data), kind="stable")]
    data = data[data.real > 0]
    data = data[np.abs(data) > 1.0e-12]
    if is_real:
        data = data.real
    return unitcell * (1 / (2 * math.pi * v)) * np.log(data[0] / data)


def central_charge(scheme, n):
    """cft.jl:256-260: M[-1;-2] := (T / n)[1 -1;-2 1], c = 6/pi log(sigma_max(M));
    BTRG (cft.jl:262-269): M := T[1 -1;3 2] S1[3;-2] S2[2;1] / n."""
    if isinstance(scheme, BTRG):
        M = np.einsum("abcd,cy,da->by", scheme.T, scheme.S1, scheme.S2, optimize=_OPT) / n
    else:
        M = np.einsum("abca->bc", scheme.T) / n
    return math.log(np.linalg.svd(M, compute_uv=False)[0]) * 6 / math.pi


def ground_state_degeneracy(scheme, unitcell=1):
    """cft.jl:278-309 / 311-339: exp of the Shannon entropy of the normalised transfer-matrix
    spectrum."""
    D = np.linalg.eigvals(transfer_matrix(scheme, unitcell))
    D = D / np.sum(D)
    vals = np.abs(D)
    vals = vals[vals > 0]
    return float(np.exp(-np.sum(vals * np.log(vals))))


def gu_wen_ratio(scheme):
    """cft.jl:374-383 / 385-395: X1 = |T[1 2;2 1]|^2 / |T[1 2;2 3] T[3 4;4 1]|,
    X2 = |T[1 2;2 1]|^2 / |T[1 2;3 4] T[4 3;2 1]|."""
    T = _unit_tensor(scheme)
    one = abs(np.einsum("abba->", T))
    x1 = abs(np.einsum("abbc,cdda->", T, T))
    x2 = abs(np.einsum("abcd,dcba->", T, T))
    return one ** 2 / x1, one ** 2 / x2


def finalize_central_charge(scheme):
    """finalize.jl:143-146."""
    n = scheme.finalize()
    return central_charge(scheme, n)


def finalize_groundstatedegeneracy(scheme):
    """finalize.jl:153-161."""
    scheme.finalize()
    return ground_state_degeneracy(scheme, 1)


def finalize_gu_wen_ratio(scheme):
    """finalize.jl:171-179."""
    scheme.finalize()
    return gu_wen_ratio(scheme)
