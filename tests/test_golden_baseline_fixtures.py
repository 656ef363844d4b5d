"""CPU checks of tests/golden/baseline_sizes.json (the fixtures of tests/test_gpu_baseline_sizes.py):
every case the device suite uses is present, was produced at the size it claims, carries its
conditioning record, and its small-size sibling is reproduced by the oracle live."""
import json
import os

import numpy as np

import tnr_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
# the cases every checkout must carry (configs[1], one configs[2] case, the 3D schemes); further
# chi = 128 cases (BTRG Z2, TRG / BTRG Z3 Potts, ATRG_3D chi = 16) are checked when present
REQUIRED = ["HOTRG_ising_trivial_chi64_it4", "TRG_ising_z2_chi128_it4",
            "HOTRG_3D_ising_trivial_chi10_it6", "HOTRG_3D_ising_trivial_chi12_it6",
            "ATRG_3D_ising_trivial_chi10_it5"]


def _gold():
    import glob

    gold = {}
    for p in sorted(glob.glob(os.path.join(HERE, "golden", "baseline_sizes*.json"))):
        with open(p) as f:      # the generator may be run in several parts (--out): merged here
            gold.update(json.load(f))
    return gold


def test_fixture_file_is_complete_and_well_conditioned():
    gold = _gold()
    missing = [n for n in REQUIRED if n not in gold]
    assert not missing, f"run tests/golden/make_golden_baseline_sizes.py {' '.join(missing)}"
    for name in gold:
        g = gold[name]
        assert g["valid"] and g["sensitivity_1e-14"] <= 1e-11, name
        assert len(g["norms"]) == g["n"] + 1 and all(np.isfinite(g["norms"])), name
        if "chi128" in name:
            assert g["dims"] == [128] * 4 and len(g["spectra"]) == 2, name
            for one in g["spectra"]:
                assert sum(len(v) for _, v in one) == 128, name      # sector-global truncrank
        if "chi64" in name:
            assert g["dims"] == [64] * 4, name


def test_first_norms_of_the_fixtures_are_the_live_oracle(tk):
    """The first three RG steps are cheap: the recorded lists start with what the oracle gives
    now (guards against a stale or hand-edited fixture)."""
    gold = _gold()
    live = o.run(o.HOTRG(np.asarray(tk.classical_ising(tk.Trivial))), 64, 2)
    assert np.allclose(gold["HOTRG_ising_trivial_chi64_it4"]["norms"][:3], live, rtol=1e-12)
    live = o.run(o.HOTRG_3D(np.asarray(tk.classical_ising_3D(tk.Trivial))), 10, 2)
    assert np.allclose(gold["HOTRG_3D_ising_trivial_chi10_it6"]["norms"][:3], live, rtol=1e-12)
    import sym_oracle as so

    T = tk.classical_ising()
    live = o.run(so.TRG_sym(np.asarray(T), T.charges, T.signs, 2), 128, 2)
    assert np.allclose(gold["TRG_ising_z2_chi128_it4"]["norms"][:3], live, rtol=1e-12)


def test_factored_cpu_vectors_agree_with_the_oracle_where_it_has_run():
    """tests/golden/factored_cpu_norms.json (ATRG_3D chi = 16 / 24, the factored step over LAPACK:
    the product's own host sequencing, VERDICT r01 "half self-referential") against the dense numpy
    ORACLE's runs at the same sizes, for every size at which the oracle fixture exists (chi = 16:
    committed; chi = 24: ~1 h of host time per run, generated in the background of round 2).
    With this, `device == factored CPU vector` (1e-10, tests/test_gpu_atrg3d_factored.py) is a
    statement about the oracle."""
    gold = _gold()
    with open(os.path.join(HERE, "golden", "factored_cpu_norms.json")) as f:
        fact = json.load(f)
    checked = 0
    for chi in (16, 24):
        ref = np.array(fact[f"ATRG_3D_ising_trivial_chi{chi}_it3"])
        for name in (f"ATRG_3D_ising_trivial_chi{chi}_it4", f"ATRG_3D_ising_trivial_chi{chi}_it3"):
            if name in gold:
                assert gold[name]["valid"], name
                oracle = np.array(gold[name]["norms"][: len(ref)])
                assert np.max(np.abs(oracle - ref) / np.abs(oracle)) <= 1e-11, (name, oracle, ref)
                checked += 1
    assert checked >= 1      # chi = 16 is committed
