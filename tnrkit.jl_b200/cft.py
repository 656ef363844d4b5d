"""Observables of the coarse-grained tensor and their finalizers for the 2D schemes
(TRG, BTRG, HOTRG, ATRG): `cft_data`, `central_charge`, `ground_state_degeneracy`,
`gu_wen_ratio` (src/utility/cft.jl:5-73, 256-339, 374-395) and `finalize_central_charge!`,
`finalize_groundstatedegeneracy!`, `finalize_gu_wen_ratio!` with their `Finalizer`s
(src/utility/finalize.jl:143-188).

Everything that scales with a power of chi beyond 2 -- the ring contraction of `unitcell`
tensors into the transfer matrix (chi^(2u+2)), the partial traces, the two-tensor norms of the
Gu-Wen ratios, the largest singular value for the central charge -- runs on the device through
the C ABI (`tnr_contract`, `tnr_axis_scale`, `tnr_svd_trunc`).  The reference then calls a
general (non-hermitian) `eig_full` on the chi^u x chi^u transfer matrix; that spectrum of one
small matrix per call is taken with LAPACK on the host, like the reference does -- it is
post-processing of a finished run, not part of `step!` / `finalize!`.

Block-sparse schemes stay block-sparse: the same contractions run sector by sector
(`sym_contract`, one grouped DMMA launch each) and only the chi^u x chi^u results leave the
device (eigenvalues of a block-diagonal matrix are the union of the block spectra, so the numbers
equal the dense ones).  `_dense_T` (host round trip of the whole tensor) is only the fallback for
tensors whose ring legs are not dual to each other.
"""
from __future__ import annotations

import math

import numpy as np

from .tensor import DeviceTensor, contract, svd_trunc


def _dense_T(scheme) -> DeviceTensor:
    T = scheme.T
    if getattr(scheme, "sym", False):
        T = DeviceTensor.from_numpy(T.to_dense(), 2, scheme.ctx)
    if len(T.dims) != 4:
        raise TypeError("cft observables are defined for the 2D schemes (4-leg tensors)")
    return T


def _weights(scheme, S, leg) -> DeviceTensor:
    """BTRG bond weights as one dense vector (block-sparse: sectors in leg order)."""
    if getattr(scheme, "sym", False):
        l = scheme.T.legs[leg]
        return DeviceTensor.from_numpy(np.concatenate([S[q].to_numpy().reshape(-1)
                                                       for q in l.charges]), 1, scheme.ctx)
    return S


def _unit_tensor(scheme) -> DeviceTensor:
    """BTRG: T_unit[-1 -2;-3 -4] := T[1 2;-3 -4] S1[-2;2] S2[-1;1] (cft.jl:44-45, 319-320,
    386-387; S1, S2 are diagonal); every other scheme: T."""
    T = _dense_T(scheme)
    if not hasattr(scheme, "S1"):
        return T
    T = T.clone()
    d = T.dims
    s2, s1 = _weights(scheme, scheme.S2, 0), _weights(scheme, scheme.S1, 1)
    T.ctx.call("tnr_axis_scale", T.ptr, 1, d[0], d[1] * d[2] * d[3], s2.ptr, 0, 0.0)
    T.ctx.call("tnr_axis_scale", T.ptr, d[0], d[1], d[2] * d[3], s1.ptr, 0, 0.0)
    return T


def _eye(n, ctx) -> DeviceTensor:
    return DeviceTensor.from_numpy(np.eye(n), 1, ctx)


def _scalar(t: DeviceTensor) -> float:
    return float(t.to_numpy().reshape(-1)[0])


# ---- block-sparse forms -------------------------------------------------------------------
def _sym_ok(scheme) -> bool:
    """Block-sparse scheme whose legs 1/4 and 2/3 are dual to each other (always the case for the
    tensors TRG / BTRG / HOTRG / ATRG produce)."""
    if not getattr(scheme, "sym", False):
        return False
    L = scheme.T.legs
    return len(L) == 4 and all(L[i].sign == -L[j].sign and L[i].same_space(L[j])
                               for i, j in ((0, 3), (1, 2)))


def _sym_eye(T, i, j):
    """delta on the legs (i, j) of T as a SymTensor that contracts with both."""
    from .symmetric import Leg, SymTensor

    li, lj = T.legs[i], T.legs[j]
    legs = [Leg(li.dims, -li.sign), Leg(lj.dims, -lj.sign)]
    blocks = {(q, q): DeviceTensor.from_numpy(np.eye(li.dims[q]), 1, T.ctx) for q in li.charges}
    return SymTensor(T.N, legs, blocks, T.ctx)


def _sym_unit_tensor(scheme):
    T = scheme.T
    if hasattr(scheme, "S1"):
        from .symmetric import sym_clone

        T = sym_clone(T).scale_leg(0, scheme.S2).scale_leg(1, scheme.S1)
    return T


def _sym_transfer_matrix(scheme, unitcell):
    from .symmetric import sym_contract

    T = _sym_unit_tensor(scheme)
    rows, cols = "bcdefg"[:unitcell], "hijklm"[:unitcell]
    R, lab = T, "a" + rows[0] + cols[0] + "z"
    for i in range(1, unitcell):
        new = "a" + rows[: i + 1] + cols[: i + 1] + "z"
        R = sym_contract(R, lab.replace("z", "y"), T, "y" + rows[i] + cols[i] + "z", new)
        lab = new
    M = sym_contract(R, lab.replace("z", "y"), _sym_eye(T, 0, 3), "ay", rows + cols).to_dense()
    n = math.prod(M.shape[:unitcell])
    return M.reshape((n, -1), order="F")


def transfer_matrix(scheme, unitcell: int = 1) -> np.ndarray:
    """ncon(fill(T, u), [[i, -i, -(i+u), i+1]..., last leg 4 -> 1]) as a matrix from the legs 2
    to the legs 3 (cft.jl:7-16, 280-295): a ring of u tensors along legs 1 / 4, contracted on the
    device; the chi^u x chi^u result is returned to the host."""
    if unitcell < 1 or unitcell > 6:
        raise ValueError("unitcell must be 1..6")
    if _sym_ok(scheme):
        return _sym_transfer_matrix(scheme, unitcell)
    T = _unit_tensor(scheme)
    rows = "bcdefg"[:unitcell]
    cols = "hijklm"[:unitcell]
    R, lab = T, "a" + rows[0] + cols[0] + "z"
    for i in range(1, unitcell):
        # running leg z -> leg 1 of the next tensor; keep rows / columns grouped
        new = "a" + rows[: i + 1] + cols[: i + 1] + "z"
        R = contract(R, lab.replace("z", "y"), T, "y" + rows[i] + cols[i] + "z", new)
        lab = new
    M = contract(R, lab.replace("z", "y"), _eye(T.dims[0], T.ctx), "ay", rows + cols)
    n = math.prod(M.dims[:unitcell])
    return M.to_numpy().reshape((n, -1), order="F")


def cft_data(scheme, v=1, unitcell=1, is_real=True):
    """cft_data(scheme; v, unitcell, is_real) -- cft.jl:5-37 (BTRG: 39-73): scaling dimensions
    from the transfer-matrix spectrum, sorted by magnitude, negative-real and < 1e-12 entries
    dropped.  The first entry is 0 (the reference's tests use `[2:end]`)."""
    data = np.linalg.eigvals(transfer_matrix(scheme, unitcell)).astype(complex)
    data = data[np.argsort(-np.abs(data), kind="stable")]
    data = data[data.real > 0]
    data = data[np.abs(data) > 1.0e-12]
    if is_real:
        data = data.real
    return unitcell * (1 / (2 * math.pi * v)) * np.log(data[0] / data)


def central_charge(scheme, n):
    """central_charge(scheme, n) -- cft.jl:256-260: M[-1;-2] := (T / n)[1 -1;-2 1],
    c = 6/pi log(sigma_max(M)); BTRG (cft.jl:262-269): M := T[1 -1;3 2] S1[3;-2] S2[2;1] / n.
    The largest singular value comes from the device SVD (`tnr_svd_trunc`, truncrank(1))."""
    if _sym_ok(scheme):
        from .symmetric import sym_clone, sym_contract, sym_svd_trunc

        T = scheme.T
        if hasattr(scheme, "S1"):
            T = sym_clone(T).scale_leg(0, scheme.S2).scale_leg(2, scheme.S1)
        M = sym_contract(T, "abcd", _sym_eye(T, 0, 3), "ad", "bc")
        _, S, _, _ = sym_svd_trunc(M, 1, 1)
        (s1,) = [float(v.to_numpy().reshape(-1)[0]) for v in S.values()]
        return math.log(s1 / abs(n)) * 6 / math.pi
    T = _dense_T(scheme)
    if hasattr(scheme, "S1"):
        d = T.dims
        s2, s1 = _weights(scheme, scheme.S2, 0), _weights(scheme, scheme.S1, 1)
        W = T.clone()                      # weight the traced leg 1 by S2 and leg 3 by S1
        W.ctx.call("tnr_axis_scale", W.ptr, 1, d[0], d[1] * d[2] * d[3], s2.ptr, 0, 0.0)
        W.ctx.call("tnr_axis_scale", W.ptr, d[0] * d[1], d[2], d[3], s1.ptr, 0, 0.0)
        T = W
    M = contract(T, "abcd", _eye(T.dims[0], T.ctx), "ad", "bc")
    _, S, _, _ = svd_trunc(M, 1, 1)
    return math.log(_scalar(S) / abs(n)) * 6 / math.pi


def ground_state_degeneracy(scheme, unitcell: int = 1):
    """ground_state_degeneracy(scheme[, unitcell]) -- cft.jl:278-309 (BTRG: 311-339): exp of the
    Shannon entropy of the transfer-matrix eigenvalues normalised by their sum."""
    D = np.linalg.eigvals(transfer_matrix(scheme, unitcell))
    vals = np.abs(D / np.sum(D))
    vals = vals[vals > 0]
    return float(np.exp(-np.sum(vals * np.log(vals))))


def gu_wen_ratio(scheme):
    """gu_wen_ratio(scheme) -- cft.jl:374-383 (BTRG: 385-395):
    X1 = |T[1 2;2 1]|^2 / |T[1 2;2 3] T[3 4;4 1]|,  X2 = |T[1 2;2 1]|^2 / |T[1 2;3 4] T[4 3;2 1]|."""
    if _sym_ok(scheme):
        from .symmetric import sym_contract

        T = _sym_unit_tensor(scheme)
        M = sym_contract(T, "abcd", _sym_eye(T, 1, 2), "bc", "ad")    # T[a 2;2 d]
        one = abs(np.trace(M.to_dense()))
        x1 = abs(np.trace(sym_contract(M, "ac", M, "cb", "ab").to_dense()))
        x2 = abs(np.trace(sym_contract(T, "abcd", T, "dcbe", "ae").to_dense()))
        return one ** 2 / x1, one ** 2 / x2
    T = _unit_tensor(scheme)
    M = contract(T, "abcd", _eye(T.dims[1], T.ctx), "bc", "ad")       # T[a 2;2 d]
    one = abs(_scalar(contract(M, "ad", _eye(T.dims[0], T.ctx), "ad", "")))
    x1 = abs(_scalar(contract(M, "ac", M, "ca", "")))
    x2 = abs(_scalar(contract(T, "abcd", T, "dcba", "")))
    return one ** 2 / x1, one ** 2 / x2


# ---- finalizers (src/utility/finalize.jl:143-188) ----------------------------------------
def finalize_central_charge(scheme):
    """finalize_central_charge!(scheme) -- finalize.jl:143-146."""
    n = scheme.finalize()
    return central_charge(scheme, n)


def finalize_groundstatedegeneracy(scheme):
    """finalize_groundstatedegeneracy!(scheme) -- finalize.jl:153-161."""
    scheme.finalize()
    return ground_state_degeneracy(scheme, 1)


def finalize_gu_wen_ratio(scheme):
    """finalize_gu_wen_ratio!(scheme) -- finalize.jl:171-179."""
    scheme.finalize()
    return gu_wen_ratio(scheme)


def _finalizers():
    from .schemes import Finalizer

    return (Finalizer(finalize_groundstatedegeneracy, float),
            Finalizer(finalize_gu_wen_ratio, tuple),
            Finalizer(finalize_central_charge, float))


# finalize.jl:163, 188 (+ the central-charge finalizer the reference's TODO at :148 asks for)
GSDegeneracy_Finalizer, guwenratio_Finalizer, central_charge_Finalizer = _finalizers()
