"""Conditioning guard for parity cases  --  TEST INFRASTRUCTURE ONLY.

`truncrank(chi)` through an (almost) degenerate multiplet of the spectrum is decided by rounding
noise in ANY implementation, the reference included (ATRG_3D chi=8 on the 3D Ising
tensor moves by 3e-2 when the input moves by 1e-14).  Such a (scheme, chi, n) is not a parity
case.  Every norm-list comparison at 1e-10 therefore first asks the ORACLE how far its own norm
list moves under a 1e-14 relative perturbation of the input tensor; a case that moves by more
than 1e-11 is refused (the test fails with a message saying the case is invalid, it does not
quietly pass)."""
from __future__ import annotations

import functools

import numpy as np

PERTURB = 1e-14
LIMIT = 1e-11


def oracle_sensitivity(run_oracle, T, seed=1234):
    """max relative change of `run_oracle(T)` (a norm list) under T -> T (1 + 1e-14 N(0,1))."""
    T = np.asarray(T, dtype=float)
    rng = np.random.default_rng(seed)
    base = np.asarray(run_oracle(T.copy()), dtype=float)
    pert = np.asarray(run_oracle(T * (1.0 + PERTURB * rng.standard_normal(T.shape))), dtype=float)
    return float(np.max(np.abs(base - pert) / np.abs(base))), base


def require_well_conditioned(run_oracle, T, what=""):
    """Returns the oracle's norm list for T; raises if the case is ill conditioned."""
    sens, base = oracle_sensitivity(run_oracle, T)
    if not sens <= LIMIT:
        raise AssertionError(
            f"invalid parity case {what}: the oracle's own norm list moves by {sens:.1e} under a "
            f"{PERTURB:g} perturbation of the input (limit {LIMIT:g}); truncrank cuts a degenerate "
            f"multiplet -- pick another chi")
    return base


@functools.lru_cache(maxsize=None)
def checked_oracle_norms(scheme, chi, n, model="ising_3d"):
    """Cached oracle norm list of a named (scheme, chi, n, model) case, conditioning-checked."""
    import tnr_oracle as o
    import tnrkit.jl_b200 as tk

    T = {"ising_3d": lambda: tk.classical_ising_3D(tk.Trivial),
         "ising_2d": lambda: tk.classical_ising(tk.Trivial)}[model]()
    cls = getattr(o, scheme)
    return tuple(require_well_conditioned(lambda t: o.run(cls(t), chi, n), np.asarray(T),
                                          f"{scheme} chi={chi} n={n} {model}"))
