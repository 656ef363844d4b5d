"""CPU tests of the block-sparse HOST SEQUENCING (tnrkit.jl_b200/symmetric.py) with the C-ABI
primitives executed by tests/abi_emulator.py instead of libtnrcuda.  What is checked here is
the sector bookkeeping -- which blocks exist, arrows, offsets, leg orders, sector-global
truncation, chunking -- against the dense / sector-aware oracle; the arithmetic of the real
primitives is checked by the `-m gpu` suite, which runs the same sequences on the device."""
import itertools

import numpy as np
import pytest

import tnr_oracle as o
from abi_emulator import EmulatedContext

RTOL = 1e-10


@pytest.fixture()
def emu(tk, monkeypatch):
    """Routes every `ctx.call` of the host layer to the numpy emulation (tests only)."""
    from tnrkit.jl_b200 import _lib

    ctx = EmulatedContext()
    monkeypatch.setattr(_lib, "_default_ctx", ctx)
    return ctx


def _random_sym(rng, N, legs):
    dims = tuple(l.total for l in legs)
    a = rng.standard_normal(dims)
    q = [np.concatenate([[c] * l.dims[c] for c in l.charges]) for l in legs]
    for idx in itertools.product(*[range(d) for d in dims]):
        if sum(l.sign * q[i][j] for i, (l, j) in enumerate(zip(legs, idx))) % N != 0:
            a[idx] = 0.0
    return a


@pytest.mark.parametrize("N", [2, 3])
def test_emulated_contract_and_svd(tk, emu, N):
    rng = np.random.default_rng(N)
    L = lambda s: tk.Leg({c: 2 + c for c in range(N)}, s)
    la = [L(+1), L(+1), L(-1), L(-1)]
    a = _random_sym(rng, N, la)
    A = tk.SymTensor.from_dense(a, N, la)
    assert np.array_equal(A.to_dense(), a)
    assert np.array_equal(A.permute((2, 0, 3, 1)).to_dense(), np.transpose(a, (2, 0, 3, 1)))
    lb = [L(+1), L(+1), L(-1)]
    b = _random_sym(rng, N, lb)
    B = tk.SymTensor.from_dense(b, N, lb)
    Cc = tk.sym_contract(A, "abxy", B, "xyc", "cab")
    assert emu.calls["tnr_gemm_grouped"] == 1
    assert np.abs(Cc.to_dense() - np.einsum("abxy,xyc->cab", a, b)).max() <= 1e-12
    U, S, V, eps = tk.sym_svd_trunc(A, 2, 5)
    n = la[0].total
    sref = np.linalg.svd(a.reshape(n * n, n * n), compute_uv=False)
    got = np.sort(np.concatenate([s.to_numpy() for s in S.values()]))[::-1]
    assert np.abs(got - sref[:5]).max() <= 1e-12 * sref[0]
    assert abs(eps - np.linalg.norm(sref[5:])) <= 1e-11 * sref[0]


@pytest.mark.parametrize("name,chi,n", [("TRG", 8, 6), ("BTRG", 8, 6), ("HOTRG", 6, 4), ("ATRG", 8, 4)])
@pytest.mark.parametrize("model", ["ising_z2", "potts_z3"])
def test_emulated_2d_block_sparse_schemes_match_oracle(tk, emu, name, chi, n, model):
    """The sequences the GPU suite verifies on the device reproduce the oracle on the emulation
    too: pins the emulation itself (tests/test_gpu_symmetric.py is the device twin)."""
    T = tk.classical_ising() if model == "ising_z2" else tk.classical_potts(3)
    s = getattr(tk, name)(T)
    assert s.sym and s.ctx is emu
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert s.T.nnz() < np.prod(s.T.dims)
