"""Block-sparse HOTRG_3D / ATRG_3D on the Z2 tensor of the reference's own 3D testsets
(`T_3D = classical_ising_3D()`, test/schemes.jl:8,365-383) through the C ABI, against the
oracle and against the dense device path.  CPU twin (same sequences on the numpy emulation of
the primitives): tests/test_host_sequencing_emulated.py."""
import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.mark.parametrize("name,chi,n", [("HOTRG_3D", 4, 3), ("HOTRG_3D", 6, 2), ("ATRG_3D", 4, 3),
                                        ("ATRG_3D", 6, 2)])
def test_block_sparse_3d_schemes_match_oracle_and_dense(tk, name, chi, n):
    T = tk.classical_ising_3D()
    cls = getattr(tk, name)
    s = cls(T, symmetric=True)
    assert s.sym
    ctx = tk.default_context()
    before = ctx.counters()["grouped_gemm_launches"]
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert ctx.counters()["grouped_gemm_launches"] > before      # per-sector grouped DMMA launches
    ref = np.array(o.run(getattr(o, name)(np.asarray(T)), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    dense = np.array(tk.run(cls(T), tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert np.max(np.abs(got - dense) / np.abs(dense)) <= RTOL
    assert s.T.nnz() < np.prod(s.T.dims)                          # only allowed blocks are stored
    f = tk.free_energy(got, tk.ising_βc_3D, scalefactor=8.0)
    fr = o.free_energy(ref, o.ising_bc_3D, scalefactor=8.0)
    assert abs(f - fr) <= RTOL * abs(fr)


def test_block_sparse_hotrg3d_chunking_is_exact(tk):
    from tnrkit.jl_b200 import symmetric

    T = tk.classical_ising_3D()
    chi, n = 6, 3
    base = np.array(tk.run(tk.HOTRG_3D(T, symmetric=True), tk.truncrank(chi), tk.maxiter(n),
                           verbosity=0))
    assert symmetric.LAST_PLAN["hotrg3d"]["chunks"] == 1
    for budget in (10, 6 ** 6 * 2):
        s = tk.HOTRG_3D(T, symmetric=True, max_chunk_elems=budget)
        got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
        assert symmetric.LAST_PLAN["hotrg3d"]["chunks"] > 1
        assert np.max(np.abs(got - base) / np.abs(base)) <= 1e-12


def test_symmetric_flag_requires_charges(tk):
    with pytest.raises(TypeError):
        tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), symmetric=True)


@pytest.mark.parametrize("name,chi,rtol", [("HOTRG_3D", 8, 1.0e-3), ("ATRG_3D", 12, 5.0e-3)])
def test_reference_3d_testsets_block_sparse(tk, name, chi, rtol):
    """test/schemes.jl:365-383 as the reference runs them: `T_3D = classical_ising_3D()` is the
    Z2Irrep tensor, so TensorKit works block by block -- here through the block-sparse path."""
    s = getattr(tk, name)(tk.classical_ising_3D(), symmetric=True)
    data = tk.run(s, tk.truncrank(chi), tk.maxiter(25), verbosity=0)
    fs = tk.free_energy(data, tk.ising_βc_3D, scalefactor=8.0)
    assert abs(fs - o.f_benchmark3D) <= rtol * abs(o.f_benchmark3D)
