"""Device twins of the cft observables (tnrkit.jl_b200/cft.py; reference src/utility/cft.jl and
the finalizers of src/utility/finalize.jl:143-188): the reference's own testsets for TRG / BTRG /
HOTRG / ATRG at their sizes and tolerances (test/schemes.jl:18-164, on the Z2 tensor
`T = classical_ising()` as there, and on the dense `Trivial` tensor), and the oracle's restatement
evaluated on the very state the device run produced.  CPU twins:
tests/test_host_sequencing_emulated.py::test_emulated_cft_observables_*,
tests/test_oracle_golden.py::test_oracle_cft_data_reference_testsets."""
import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu


def _twin(name, s):
    dense = s.T.to_dense() if s.sym else s.T.to_numpy()
    tw = getattr(o, name)(dense)
    if name == "BTRG":
        if s.sym:
            w = lambda S, leg: np.concatenate([S[q].to_numpy().reshape(-1) for q in s.T.legs[leg].charges])
            tw.S1, tw.S2 = np.diag(w(s.S1, 1)), np.diag(w(s.S2, 0))
        else:
            tw.S1, tw.S2 = np.diag(s.S1.to_numpy().reshape(-1)), np.diag(s.S2.to_numpy().reshape(-1))
    return tw


@pytest.mark.parametrize("sector", ["Z2", "Trivial"])
@pytest.mark.parametrize("name,chi,n,r1,r2", [("TRG", 24, 10, 2.0e-4, 1.0e-2), ("BTRG", 24, 10, 3.0e-4, 2.0e-2),
                                              ("HOTRG", 16, 4, 6.0e-4, 1.0e-2), ("ATRG", 24, 3, 1.0e-2, 1.0e-2)])
def test_cft_data_reference_testsets(tk, name, chi, n, r1, r2, sector):
    T = tk.classical_ising() if sector == "Z2" else tk.classical_ising(tk.Trivial)
    s = getattr(tk, name)(T)
    tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    cft = tk.cft_data(s)[1:]
    assert abs(cft[0] - o.ising_cft_exact[0]) <= r1 * o.ising_cft_exact[0]
    assert abs(cft[1] - o.ising_cft_exact[1]) <= r2 * o.ising_cft_exact[1]
    tw = _twin(name, s)
    assert np.abs(cft[:6] - o.cft_data(tw)[1:7]).max() <= 1e-8
    assert abs(tk.central_charge(s, 1.3) - o.central_charge(tw, 1.3)) <= 1e-10
    assert np.abs(tk.cft_data(s, unitcell=2)[:6] - o.cft_data(tw, unitcell=2)[:6]).max() <= 1e-7


@pytest.mark.parametrize("name,chi", [("TRG", 16), ("BTRG", 16), ("HOTRG", 12), ("ATRG", 16)])
@pytest.mark.parametrize("dbeta,want", [(-0.01, 1.0), (+0.01, 2.0)])
def test_gsd_and_gu_wen_reference_testsets(tk, name, chi, dbeta, want):
    s = getattr(tk, name)(tk.classical_ising(o.ising_bc + dbeta))
    tk.run(s, tk.truncrank(chi), tk.maxiter(20), verbosity=0)
    x1, x2 = tk.gu_wen_ratio(s)
    gsd = tk.ground_state_degeneracy(s)
    for got in (gsd, x1, x2):
        assert abs(got - want) <= 1.0e-2 * want
    tw = _twin(name, s)
    assert abs(gsd - o.ground_state_degeneracy(tw)) <= 1e-10
    assert np.abs(np.array([x1, x2]) - np.array(o.gu_wen_ratio(tw))).max() <= 1e-10
    assert abs(tk.finalize_groundstatedegeneracy(s) - want) <= 1.0e-2 * want


def test_observable_finalizers_through_run(tk):
    s = tk.TRG(tk.classical_ising(tk.Trivial))
    data = tk.run(s, tk.truncrank(8), tk.maxiter(4), tk.guwenratio_Finalizer, verbosity=0)
    tw = o.TRG(o.classical_ising())
    ref = [o.finalize_gu_wen_ratio(tw)]
    for _ in range(4):
        tw.step(8)
        ref.append(o.finalize_gu_wen_ratio(tw))
    assert np.abs(np.array(data) - np.array(ref)).max() <= 1e-9
    data = tk.run(tk.HOTRG(tk.classical_ising(tk.Trivial)), tk.truncrank(8), tk.maxiter(3),
                  tk.central_charge_Finalizer, verbosity=0)
    tw = o.HOTRG(o.classical_ising())
    ref = [o.finalize_central_charge(tw)]
    for _ in range(3):
        tw.step(8)
        ref.append(o.finalize_central_charge(tw))
    assert np.abs(np.array(data) - np.array(ref)).max() <= 1e-9


@pytest.mark.parametrize("sector", ["Z2", "Trivial"])
@pytest.mark.parametrize("name", ["TRG", "BTRG", "HOTRG", "ATRG"])
def test_two_by_two_finalizer_all_schemes(tk, name, sector):
    """finalize_two_by_two! (finalize.jl:17-25; BTRG with bond weights: 27-42) through run!."""
    T = tk.classical_ising() if sector == "Z2" else tk.classical_ising(tk.Trivial)
    got = tk.run(getattr(tk, name)(T), tk.truncrank(8), tk.maxiter(4), tk.two_by_two_Finalizer,
                 verbosity=0)
    tw = getattr(o, name)(o.classical_ising())
    ref = [o.finalize_two_by_two(tw)]
    for _ in range(4):
        tw.step(8)
        ref.append(o.finalize_two_by_two(tw))
    assert np.max(np.abs(np.array(got) - np.array(ref)) / np.abs(ref)) <= 1e-10
