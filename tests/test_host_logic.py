"""CPU tests of the host-side mirror: stopping criteria, truncation strategies, model
constructors, sharding helpers, C-ABI symbol export, loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import tnr_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stopping_criteria(tk):
    assert tk.maxiter(3)(2, []) and not tk.maxiter(3)(3, [])
    c = tk.convcrit(1e-3, lambda steps, data: abs(data[-1]))
    assert c(1, [1.0]) and not c(1, [1e-4])
    both = tk.maxiter(5) & c
    assert both(1, [1.0]) and not both(5, [1.0]) and not both(1, [1e-9])
    assert "Maximum" in both.info(5, [1.0]) and "Convergence" in both.info(1, [1e-9])
    assert not tk.trivial_convcrit(1e-2)(1, [1e-3])


def test_truncation_strategies(tk):
    assert tk.truncrank(16).chi == 16
    with pytest.raises(ValueError):
        tk.truncrank(0)
    with pytest.raises(NotImplementedError):
        tk.trunctol(atol=1e-10)
    with pytest.raises(NotImplementedError):
        tk.truncrank(4) & tk.truncrank(5)


def test_models_match_oracle(tk):
    assert np.array_equal(tk.classical_ising(), o.classical_ising_z2basis())
    assert np.array_equal(tk.classical_ising(tk.Z2Irrep, 0.3), o.classical_ising_z2basis(0.3))
    assert np.array_equal(tk.classical_ising(tk.Trivial, 0.3, h=0.2), o.classical_ising(0.3, 0.2))
    assert np.array_equal(tk.classical_ising_3D(tk.Trivial), o.classical_ising_3D())
    assert np.array_equal(tk.classical_ising_3D(0.2), o.classical_ising_3D_z2basis(0.2))
    assert np.array_equal(tk.classical_potts(tk.Trivial, 3), o.classical_potts(3))
    with pytest.raises(AssertionError):
        tk.classical_ising(tk.Z2Irrep, 0.3, h=0.1)
    with pytest.raises(AssertionError):
        tk.classical_potts(tk.ZNIrrep[4], 3)


def test_potts_zn_basis_is_block_sparse(tk):
    q = 3
    t = tk.classical_potts(q)  # ZNIrrep{3}: charge conservation i+j = k+l mod q
    for i, j, k, l in np.ndindex(q, q, q, q):
        if (i + j - k - l) % q != 0:
            assert abs(t[i, j, k, l]) < 1e-13
    # same network as the Trivial tensor up to a unitary gauge: same TRG free energy
    f_sym = o.free_energy(o.run(o.TRG(t), 9, 8), o.potts_bc(3))
    f_triv = o.free_energy(o.run(o.TRG(o.classical_potts(3)), 9, 8), o.potts_bc(3))
    assert abs(f_sym - f_triv) < 1e-9 * abs(f_triv)


def test_free_energy_matches_oracle(tk):
    data = [1.3, 0.7, 2.1, 1.01]
    for kw in ({}, {"scalefactor": 4.0}, {"scalefactor": 8.0, "initial_size": 2.0}):
        assert tk.free_energy(data, 0.44, **kw) == o.free_energy(data, 0.44, **kw)


def test_shard_range_partitions(tk):
    for n in (1, 3, 24, 25):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = tk.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi))
            assert cover == list(range(n))
            sizes = [hi - lo for lo, hi in (tk.shard_range(n, r, world) for r in range(world))]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_abi_exports_every_declared_symbol(tk):
    from tnrkit.jl_b200 import _lib

    header = open(os.path.join(ROOT, "include", "tnrcuda.h")).read()
    declared = set(re.findall(r"\b(tnr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tnr_version() == 100


def test_no_cpu_fallback(tk):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tk.TNRCudaError):
        tk.TRG(tk.classical_ising())
    from tnrkit.jl_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.tnr_create(0, None, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.tnr_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tnrkit.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "tnr_oracle" not in src and "import oracle" not in src, f


def test_symmetric_sector_bookkeeping(tk):
    """Host-side structure of block-sparse Z_N tensors (no device needed): allowed charge
    tuples, coupled-sector row/column offsets."""
    L = lambda s: tk.Leg({0: 2, 1: 3, 2: 1}, s)
    t = tk.SymTensor(3, [L(+1), L(+1), L(-1), L(-1)], {})
    keys = list(t.keys())
    assert len(keys) == 27 and all((a + b - c - d) % 3 == 0 for a, b, c, d in keys)
    assert t.dims == (6, 6, 6, 6)
    assert t.block_dims((0, 1, 1, 0)) == (2, 3, 3, 2)
    rows = t._tuples([0, 1], negate=False)
    cols = t._tuples([2, 3], negate=True)
    assert sorted(rows) == sorted(cols) == [0, 1, 2]
    for c in (0, 1, 2):
        # offsets are cumulative sizes; rows and columns of a coupled sector have equal totals
        # here because both pairs of legs carry the same spaces
        tot_r = rows[c][-1][1] + rows[c][-1][2]
        tot_c = cols[c][-1][1] + cols[c][-1][2]
        assert tot_r == tot_c == sum(t.legs[0].dims[a] * t.legs[1].dims[b]
                                     for a in range(3) for b in range(3) if (a + b) % 3 == c)
        off = 0
        for key, o_, size in rows[c]:
            assert o_ == off and sum(key) % 3 == c
            off += size
    leg = tk.Leg({2: 1, 0: 4}, -1)
    assert leg.charges == (0, 2) and leg.offsets == {0: 0, 2: 4} and leg.total == 5
    assert leg.flipped().sign == +1 and leg.same_space(leg.flipped())


def test_charged_model_arrays(tk):
    t = tk.classical_ising()
    assert isinstance(t, tk.ChargedArray) and t.N == 2 and t.signs == (1, 1, -1, -1)
    p = tk.classical_potts(3)
    assert p.N == 3 and p.charges[0] == (0, 1, 2)
    assert type(np.asarray(p)) is np.ndarray
    assert getattr(tk.classical_ising(tk.Trivial), "charges", None) is None


def test_header_is_plain_c():
    """include/tnrcuda.h must compile as C99 (it is what a Julia `ccall` / any FFI binds)."""
    import subprocess

    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "tnrcuda.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("name,kind,dims,chi", [
    ("TRG", 0, (2, 3, 3, 2), 4), ("TRG", 0, (4, 4, 4, 4), 5), ("BTRG", 1, (2, 3, 3, 2), 4),
    ("BTRG", 1, (3, 3, 3, 3), 20), ("HOTRG", 2, (2, 3, 3, 2), 5), ("HOTRG", 2, (4, 4, 4, 4), 6),
    ("ATRG", 3, (2, 3, 3, 2), 4), ("ATRG", 3, (3, 3, 3, 3), 5), ("ATRG", 3, (2, 2, 2, 2), 16),
    ("HOTRG_3D", 4, (2, 2, 2, 2, 2, 2), 3), ("HOTRG_3D", 4, (2, 2, 3, 2, 3, 2), 5),
    ("ATRG_3D", 5, (2, 2, 2, 2, 2, 2), 3), ("ATRG_3D", 5, (2, 2, 2, 2, 2, 2), 9),
])
def test_step_out_dims_matches_oracle_shapes(tk, name, kind, dims, chi):
    """tnr_step_out_dims is pure host arithmetic: it must predict the leg dimensions the
    reference algorithm produces (checked against the oracle on a random tensor)."""
    from tnrkit.jl_b200 import _lib

    lib = _lib.load()
    out = (ctypes.c_int64 * len(dims))()
    assert lib.tnr_step_out_dims(kind, _lib.i64(dims), chi, out) == 0
    rng = np.random.default_rng(sum(dims) + chi)
    s = getattr(o, name)(rng.random(dims) + 0.05)
    s.step(chi)
    assert tuple(out) == s.T.shape
