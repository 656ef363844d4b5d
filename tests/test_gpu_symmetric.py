"""Block-sparse (Z2 / ZN) path: per-sector grouped GEMM + per-sector SVD with sector-global
truncrank, against numpy on the dense embedding and against the oracle's norm lists."""
import itertools

import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _random_sym(tk, rng, N, legs):
    dims = tuple(l.total for l in legs)
    a = rng.standard_normal(dims)
    q = [np.concatenate([[c] * l.dims[c] for c in l.charges]) for l in legs]
    for idx in itertools.product(*[range(d) for d in dims]):
        if sum(l.sign * q[i][j] for i, (l, j) in enumerate(zip(legs, idx))) % N != 0:
            a[idx] = 0.0
    return a


@pytest.mark.parametrize("N", [2, 3])
def test_symtensor_roundtrip_permute_contract(tk, N):
    rng = np.random.default_rng(N)
    L = lambda s: tk.Leg({c: 2 + c for c in range(N)}, s)
    la = [L(+1), L(+1), L(-1), L(-1)]
    a = _random_sym(tk, rng, N, la)
    A = tk.SymTensor.from_dense(a, N, la)
    assert np.array_equal(A.to_dense(), a)
    assert np.array_equal(A.permute((2, 0, 3, 1)).to_dense(), np.transpose(a, (2, 0, 3, 1)))
    lb = [L(+1), L(+1), L(-1)]
    b = _random_sym(tk, rng, N, lb)
    B = tk.SymTensor.from_dense(b, N, lb)
    ctx = tk.default_context()
    before = ctx.counters()["grouped_gemm_launches"]
    # contract A legs (2,3) [signs -] with B legs (0,1) [signs +]
    Cc = tk.sym_contract(A, "abxy", B, "xyc", "cab")
    assert ctx.counters()["grouped_gemm_launches"] == before + 1  # all sectors in ONE launch
    ref = np.einsum("abxy,xyc->cab", a, b)
    assert np.abs(Cc.to_dense() - ref).max() <= 1e-12
    with pytest.raises(ValueError):
        bad = a.copy()
        bad[0, 0, 0, 1] = 1.0 if (0 + 0 - 0 - 1) % N != 0 or N == 1 else bad[0, 0, 0, 1]
        bad[(0, 0, 0, la[3].total - 1)] = 1.0
        tk.SymTensor.from_dense(bad, N, la)


@pytest.mark.parametrize("N,chi", [(2, 5), (3, 7), (3, 100)])
def test_sym_svd_global_truncation(tk, N, chi):
    rng = np.random.default_rng(10 * N + chi)
    L = lambda s: tk.Leg({c: 3 + (c % 2) for c in range(N)}, s)
    legs = [L(+1), L(+1), L(-1), L(-1)]
    a = _random_sym(tk, rng, N, legs)
    T = tk.SymTensor.from_dense(a, N, legs)
    U, S, V, eps = tk.sym_svd_trunc(T, 2, chi)
    n = legs[0].total
    sref = np.linalg.svd(a.reshape(n * n, n * n), compute_uv=False)
    k = min(chi, sref.size)
    got = np.sort(np.concatenate([s.to_numpy() for s in S.values()]))[::-1]
    assert got.size == k
    assert np.abs(got - sref[:k]).max() <= 1e-12 * sref[0]       # sector-global top-chi
    assert abs(eps - np.linalg.norm(sref[k:])) <= 1e-11 * sref[0]
    # reconstruct: U S V equals the best rank-k approximation of the dense matrix
    u, v = U.to_dense(), V.to_dense()
    sfull = np.concatenate([S[c].to_numpy() for c in U.legs[2].charges])
    rec = np.einsum("abk,k,kcd->abcd", u, sfull, v)
    uu, ss, vv = np.linalg.svd(a.reshape(n * n, n * n), full_matrices=False)
    best = ((uu[:, :k] * ss[:k]) @ vv[:k]).reshape(a.shape)
    assert np.abs(rec - best).max() <= 1e-10 * sref[0]


@pytest.mark.parametrize("name,chi,n", [("TRG", 8, 6), ("BTRG", 8, 6), ("HOTRG", 6, 4), ("ATRG", 8, 4)])
@pytest.mark.parametrize("model", ["ising_z2", "potts_z3"])
def test_block_sparse_schemes_match_oracle_and_dense(tk, name, chi, n, model):
    T = tk.classical_ising() if model == "ising_z2" else tk.classical_potts(3)
    cls, ocls = getattr(tk, name), getattr(o, name)
    s = cls(T)
    assert s.sym, "charged tensors must take the block-sparse path"
    ctx = tk.default_context()
    before = ctx.counters()["grouped_gemm_launches"]
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert ctx.counters()["grouped_gemm_launches"] >= before + 3 * n
    ref = np.array(o.run(ocls(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    dense = np.array(tk.run(cls(T, symmetric=False), tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert np.max(np.abs(got - dense) / np.abs(dense)) <= RTOL
    # the coarse-grained tensor stays block sparse: only symmetry-allowed blocks are stored
    assert s.T.nnz() < np.prod(s.T.dims)


@pytest.mark.parametrize("model,N", [("ising_z2", 2), ("potts_z3", 3)])
def test_retained_sector_spectra_match_sector_oracle(tk, model, N):
    """The retained singular-value spectra, sector by sector, after several block-sparse TRG steps
    against the sector-aware CPU oracle (TensorKit semantics: per-sector SVD, global truncrank)."""
    import sym_oracle as so
    from tnrkit.jl_b200 import symmetric

    T = tk.classical_ising() if model == "ising_z2" else tk.classical_potts(3)
    chi, n = 9, 5
    s_gpu = tk.TRG(T)
    got = tk.run(s_gpu, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    s_ref = so.TRG_sym(np.asarray(T), T.charges, T.signs, N)
    ref = o.run(s_ref, chi, n)
    assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= RTOL
    for S_gpu, sp_ref in zip(symmetric.LAST_SPECTRA["trg"], s_ref.last_spectra):
        assert sorted(S_gpu) == [c for c, _ in sp_ref]           # same sectors kept
        for c, vals in sp_ref:
            g = S_gpu[c].to_numpy()
            assert g.shape == vals.shape                           # same multiplicity per sector
            assert np.abs(g - vals).max() <= 1e-10 * vals.max()


@pytest.mark.parametrize("name,chi,n,sf", [("TRG", 8, 5, 2.0), ("HOTRG", 6, 3, 4.0), ("ATRG", 7, 3, 4.0)])
def test_coarse_grained_tensor_spectrum_matches_oracle(tk, name, chi, n, sf):
    """Gauge-invariant content of the coarse-grained tensor itself: singular values of T viewed as
    a (1 2 | 3 4) matrix after n steps, dense path vs oracle."""
    T = tk.classical_ising(tk.Trivial, 0.41)
    s = getattr(tk, name)(T)
    tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    r = getattr(o, name)(T)
    o.run(r, chi, n)
    a = s.T.to_numpy()
    sv_g = np.linalg.svd(a.reshape(a.shape[0] * a.shape[1], -1), compute_uv=False)
    sv_r = np.linalg.svd(r.T.reshape(r.T.shape[0] * r.T.shape[1], -1), compute_uv=False)
    k = min(chi, sv_r.size)
    assert np.abs(sv_g[:k] - sv_r[:k]).max() <= 1e-9 * sv_r[0]
