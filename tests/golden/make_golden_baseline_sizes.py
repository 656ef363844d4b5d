"""Generates tests/golden/baseline_sizes.json: oracle outputs AT THE SIZES BASELINE.json NAMES
(configs[1]: HOTRG chi=64; configs[2]: TRG / BTRG on Z2 Ising and Z3 Potts at chi=128; the 3D
schemes at the largest chi the numpy oracle can hold), for the `-m gpu` parity tests in
tests/test_gpu_baseline_sizes.py (1e-10 on the norm lists and the retained spectra).

Every case is run twice: on the model tensor and on the tensor perturbed by 1e-14 (relative,
multiplicative, so symmetry zeros stay zero).  The file records the oracle's own sensitivity; a
case whose norm list moves by more than 1e-11 is recorded with "valid": false and the GPU test
refuses it (tests/conditioning.py explains why such a case is no parity target).

The oracle is pinned to the reference's published numbers by tests/test_oracle_golden.py; the
reference is Julia and cannot run in this image.  Block-sparse cases use oracle/sym_oracle.py
(TensorKit semantics: per-sector SVD, sector-global truncrank).

    python tests/golden/make_golden_baseline_sizes.py [case ...]      # ~1 h on 8 cores
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sym_oracle as so  # noqa: E402
import tnr_oracle as o  # noqa: E402

import tnrkit.jl_b200 as tk  # noqa: E402  (model constructors only: host code, no device)

PATH = os.path.join(HERE, "baseline_sizes.json")
PERTURB = 1e-14


def spectra_of(s):
    sp = getattr(s, "last_spectra", None)
    if sp is None:
        return None
    return [[[int(c), [float(x) for x in vals]] for c, vals in one] for one in sp]


def run_case(make, chi, n):
    rng = np.random.default_rng(1234)
    s = make(None)
    t0 = time.time()
    base = o.run(s, chi, n)
    dt = time.time() - t0
    pert = o.run(make(rng), chi, n)
    sens = float(np.max(np.abs(np.array(base) - np.array(pert)) / np.abs(np.array(base))))
    return {"chi": chi, "n": n, "norms": [float(x) for x in base], "sensitivity_1e-14": sens,
            "valid": bool(sens <= 1e-11), "dims": list(s.T.shape), "spectra": spectra_of(s),
            "oracle_seconds": round(dt, 1)}


def perturbed(T, rng):
    T = np.asarray(T, dtype=float)
    return T if rng is None else T * (1.0 + PERTURB * rng.standard_normal(T.shape))


def dense(cls, T):
    return lambda rng: cls(perturbed(T, rng))


def sym(cls, T, N):
    return lambda rng: cls(perturbed(np.asarray(T), rng), T.charges, T.signs, N)


T2 = tk.classical_ising()          # Z2Irrep
P3 = tk.classical_potts(3)         # ZNIrrep{3}
CASES = {
    # configs[1]: bonds 2 -> 4 -> 16 -> 64 (from 256) -> 64 (from 4096): step 4 is full size
    "HOTRG_ising_trivial_chi64_it4": (dense(o.HOTRG, tk.classical_ising(tk.Trivial)), 64, 4),
    # configs[2]: bonds 2 -> 4 -> 16 -> 128 (from 256) -> 128 (from 16384): step 4 is full size
    "TRG_ising_z2_chi128_it4": (sym(so.TRG_sym, T2, 2), 128, 4),
    "BTRG_ising_z2_chi128_it4": (sym(so.BTRG_sym, T2, 2), 128, 4),
    # bonds 3 -> 9 -> 81 -> 128 (from 6561) -> 128 (from 16384)
    "TRG_potts_z3_chi128_it4": (sym(so.TRG_sym, P3, 3), 128, 4),
    "BTRG_potts_z3_chi128_it4": (sym(so.BTRG_sym, P3, 3), 128, 4),
    # 3D: the largest bond dimensions the dense numpy oracle holds (chi^8 doubles / 16384^2 SVDs)
    "HOTRG_3D_ising_trivial_chi10_it6": (dense(o.HOTRG_3D, tk.classical_ising_3D(tk.Trivial)), 10, 6),
    "HOTRG_3D_ising_trivial_chi12_it6": (dense(o.HOTRG_3D, tk.classical_ising_3D(tk.Trivial)), 12, 6),
    "ATRG_3D_ising_trivial_chi16_it4": (dense(o.ATRG_3D, tk.classical_ising_3D(tk.Trivial)), 16, 4),
    "ATRG_3D_ising_trivial_chi10_it5": (dense(o.ATRG_3D, tk.classical_ising_3D(tk.Trivial)), 10, 5),
    # chi = 24 (1.5 GB tensors, 13824 x 13824 SVDs: ~1 h of host time per run, ~20 GB): bonds
    # 2 -> 4 -> 16 -> 24, so step 4 is a full-size step.  Replaces the vector that the factored step
    # itself produced over LAPACK (factored_cpu_norms.json) as the reference of the chi = 24 tests.
    "ATRG_3D_ising_trivial_chi24_it4": (dense(o.ATRG_3D, tk.classical_ising_3D(tk.Trivial)), 24, 4),
    "ATRG_3D_ising_trivial_chi24_it3": (dense(o.ATRG_3D, tk.classical_ising_3D(tk.Trivial)), 24, 3),
}

if __name__ == "__main__":
    argv = sys.argv[1:]
    if "--out" in argv:          # separate output file (merge by hand): parallel generation
        i = argv.index("--out")
        PATH = argv[i + 1]
        del argv[i:i + 2]
    want = argv or list(CASES)
    out = {}
    if os.path.exists(PATH):
        with open(PATH) as f:
            out = json.load(f)
    for name in want:
        make, chi, n = CASES[name]
        t0 = time.time()
        out[name] = run_case(make, chi, n)
        print(f"{name}: valid={out[name]['valid']} sensitivity={out[name]['sensitivity_1e-14']:.1e} "
              f"{time.time() - t0:.0f}s", flush=True)
        with open(PATH, "w") as f:
            json.dump(out, f, indent=1)
