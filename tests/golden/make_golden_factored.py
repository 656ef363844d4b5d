"""Generates tests/golden/factored_cpu_norms.json: norm lists of ATRG_3D at sizes where the dense
numpy oracle is too slow (chi = 24: 1.5 GB tensors, 13824^2 SVDs), computed on the CPU by the
FACTORED step (tnrkit.jl_b200/atrg3d_factored.py) with the C-ABI primitives executed by the numpy /
LAPACK emulation tests/abi_emulator.py.  That path is pinned to the oracle at chi <= 12
(tests/test_atrg3d_factored_emulated.py); two runs with different subspace blocks (56, 72) agree to
2e-14 at chi = 24, so the vector is well conditioned.  The `-m gpu` suite compares the DENSE device
step (`tnr_atrg3d_step`, a different algorithm on different hardware) and the factored device step
with it.  Takes ~4 minutes:  python tests/golden/make_golden_factored.py
(`--chi48` adds the chi = 48 vector of BASELINE.json configs[3]: ~70 minutes.)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from abi_emulator import EmulatedContext  # noqa: E402

import tnrkit.jl_b200 as tk  # noqa: E402
from tnrkit.jl_b200 import _lib  # noqa: E402

_lib._default_ctx = EmulatedContext()
path = os.path.join(HERE, "factored_cpu_norms.json")
out = json.load(open(path)) if os.path.exists(path) else {}
if "--chi48" in sys.argv:
    # BASELINE.json configs[3] size: chi = 48 (one chi^6 tensor would be 98 GB).  R factors from the
    # Gram matrices (rfactor="gram"), chunks of 2^27 doubles; about 70 minutes on 8 host cores,
    # almost all of it in numpy's einsum for the H / G chunks
    s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), factored=True, rfactor="gram",
                   max_chunk_elems=1 << 27)
    out["ATRG_3D_ising_trivial_chi48_it3_gram"] = tk.run(s, tk.truncrank(48), tk.maxiter(3),
                                                         verbosity=0)
else:
    for chi, n in ((16, 3), (24, 3)):
        s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), factored=True)
        out[f"ATRG_3D_ising_trivial_chi{chi}_it{n}"] = tk.run(s, tk.truncrank(chi), tk.maxiter(n),
                                                             verbosity=0)
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(out)
