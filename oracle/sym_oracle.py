"""CPU oracle for the block-sparse (abelian Z_N, U(1): N = 0) semantics -- TEST INFRASTRUCTURE ONLY.

TensorKit's `svd_trunc(t; trunc = truncrank(chi))` on a `Z2Irrep` / `ZNIrrep{N}` TensorMap
decomposes every coupled-sector block separately and keeps the chi largest singular values
over ALL sectors (MatrixAlgebraKit truncation on the concatenated spectra).  This file restates
that on dense charge-basis arrays with numpy; the schemes reuse tnr_oracle's contractions.
Same import rules as tnr_oracle.py (tests / smoke / bench cpu leg only).
"""
from __future__ import annotations

import numpy as np

import tnr_oracle as o


def sector_svd_trunc(T, ncod, chi, charges, signs, N):
    """Per-sector SVD + sector-global truncrank.  `charges[leg][i]` = irrep label of index i,
    `signs[leg]` = +1 / -1.  Returns dense U (cod.., k), s (k,), Vh (k, dom..), eps and the list of
    (sector, kept spectrum) -- bond index ordered by sector, then by decreasing value."""
    cod, dom = T.shape[:ncod], T.shape[ncod:]
    m, n = int(np.prod(cod)), int(np.prod(dom))
    M = T.reshape(m, n)

    def fused(shape, legs, negate):
        q = np.zeros(shape, dtype=int)
        for ax, leg in enumerate(legs):
            sh = [1] * len(shape)
            sh[ax] = shape[ax]
            q = q + signs[leg] * np.asarray(charges[leg]).reshape(sh)
        q = -q if negate else q
        return (q % N if N else q).reshape(-1)      # N = 0: U(1), no modulus

    qr = fused(cod, range(ncod), False)
    qc = fused(dom, range(ncod, T.ndim), True)
    facs, allv = {}, []
    for c in sorted(set(qr.tolist()) | set(qc.tolist())):
        r, cidx = np.where(qr == c)[0], np.where(qc == c)[0]
        if len(r) == 0 or len(cidx) == 0:
            continue
        U, s, Vh = np.linalg.svd(M[np.ix_(r, cidx)], full_matrices=False)
        facs[c] = (r, cidx, U, s, Vh)
        allv += [(v, c, j) for j, v in enumerate(s)]
    # everything outside the sector blocks must vanish (symmetric tensor)
    mask = qr[:, None] == qc[None, :]
    assert np.abs(M[~mask]).max(initial=0.0) <= 1e-12 * max(1.0, np.abs(M).max())
    allv.sort(key=lambda t: -t[0])
    keep = allv[:chi]
    eps = float(np.sqrt(sum(v * v for v, _, _ in allv[chi:])))
    kc = {c: sum(1 for _, cc, _ in keep if cc == c) for c in facs}
    k = sum(kc.values())
    U = np.zeros((m, k))
    Vh = np.zeros((k, n))
    s = np.zeros(k)
    spectra, pos = [], 0
    for c in sorted(facs):
        r, cidx, Uc, sc, Vc = facs[c]
        kk = kc[c]
        if kk == 0:
            continue
        U[np.ix_(r, range(pos, pos + kk))] = Uc[:, :kk]
        Vh[np.ix_(range(pos, pos + kk), cidx)] = Vc[:kk]
        s[pos:pos + kk] = sc[:kk]
        spectra.append((c, sc[:kk].copy()))
        pos += kk
    bond_charges = sum([[c] * len(sp) for c, sp in spectra], [])
    return U.reshape(*cod, k), s, Vh.reshape(k, *dom), eps, spectra, bond_charges


class TRG_sym:
    """TRG (src/schemes/trg.jl) with TensorKit's symmetric svd_trunc semantics."""

    def __init__(self, T, charges, signs, N):
        self.T = np.array(T, dtype=float)
        self.charges = [list(c) for c in charges]
        self.signs = list(signs)
        self.N = N
        self.last_spectra = None

    def step(self, chi):
        T, ch, sg, N = self.T, self.charges, self.signs, self.N
        U, s, V, _, sp1, b1 = sector_svd_trunc(T, 2, chi, ch, sg, N)
        A, B = U * np.sqrt(s), np.sqrt(s)[:, None, None] * V
        perm = (1, 3, 0, 2)
        Tp = np.transpose(T, perm)
        U, s, V, _, sp2, b2 = sector_svd_trunc(Tp, 2, chi, [ch[p] for p in perm],
                                               [sg[p] for p in perm], N)
        C, D = U * np.sqrt(s), np.sqrt(s)[:, None, None] * V
        self.T = np.einsum("bpq,asp,src,rqd->abcd", D, B, C, A, optimize=o._OPT)
        # new legs: a = bond of B (+, sector c1), b = bond of D (+, c2), c = bond of C (-), d = A (-)
        self.charges = [b1, b2, b2, b1]
        self.signs = [1, 1, -1, -1]
        self.last_spectra = (sp1, sp2)

    finalize = o.TRG.finalize


class BTRG_sym:
    """BTRG (src/schemes/btrg.jl:62-97) with TensorKit's symmetric svd_trunc semantics: the two
    truncated SVDs of `step!` are done sector by sector with a sector-global truncrank; bond
    weights, contraction and finalize! are tnr_oracle.BTRG's."""

    def __init__(self, T, charges, signs, N, k=-0.5):
        self.T = np.array(T, dtype=float)
        self.charges = [list(c) for c in charges]
        self.signs = list(signs)
        self.N = N
        self.k = k
        self.S1 = np.eye(self.T.shape[1])
        self.S2 = np.eye(self.T.shape[0])
        self.last_spectra = None

    def step(self, chi):
        T, ch, sg, N, k = self.T, self.charges, self.signs, self.N, self.k
        U, S, V, _, sp1, b1 = sector_svd_trunc(T, 2, chi, ch, sg, N)
        Sa, Sb = o.pseudopow(S, (1 - k) / 2), o.pseudopow(S, k)
        A, B, S1n = U * Sa, Sa[:, None, None] * V, np.diag(Sb)
        perm = (2, 0, 3, 1)                                     # ((3,1),(4,2)), btrg.jl:75
        U, S, V, _, sp2, b2 = sector_svd_trunc(np.transpose(T, perm), 2, chi,
                                               [ch[p] for p in perm], [sg[p] for p in perm], N)
        Sa, Sb = o.pseudopow(S, (1 - k) / 2), o.pseudopow(S, k)
        C, D, S2n = U * Sa, Sa[:, None, None] * V, np.diag(Sb)
        self.T = np.einsum("aqt,it,bik,kj,ujd,uo,poc,qp->abcd",
                           D, self.S1, B, self.S2, C, self.S1, A, self.S2, optimize=o._OPT)
        self.S1, self.S2 = S1n, S2n
        # new legs: a = bond of D (V of the 2nd SVD, +), b = bond of B (V of the 1st, +),
        #           c = bond of A (U of the 1st, -), d = bond of C (U of the 2nd, -)
        self.charges = [b2, b1, b1, b2]
        self.signs = [1, 1, -1, -1]
        self.last_spectra = (sp1, sp2)

    finalize = o.BTRG.finalize
