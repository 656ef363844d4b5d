"""Secondary configurations of BASELINE.json (configs[1], configs[2]): seconds per step at
steady-state bond dimension + free-energy sanity against the exact value.  Not the headline
bench (bench.py); numbers go to profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tnrkit.jl_b200 as tk

ctx = tk.default_context()
which = sys.argv[1:] or ["hotrg64", "trg128", "btrg128z2", "trg128z3"]
cases = {
    "hotrg64": (lambda: tk.HOTRG(tk.classical_ising(tk.Trivial)), 64, 7, 4.0, tk.ising_βc, tk.f_onsager),
    "trg128": (lambda: tk.TRG(tk.classical_ising(tk.Trivial)), 128, 7, 2.0, tk.ising_βc, tk.f_onsager),
    "btrg128z2": (lambda: tk.BTRG(tk.classical_ising(tk.Z2Irrep)), 128, 7, 2.0, tk.ising_βc, tk.f_onsager),
    "trg128z3": (lambda: tk.TRG(tk.classical_potts(3)), 128, 7, 2.0, tk.potts_βc(3), -4.119552029995684),
    "atrg3d16": (lambda: tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial)), 16, 4, 8.0, tk.ising_βc_3D, -3.507),
    "atrg3d24": (lambda: tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial)), 24, 4, 8.0, tk.ising_βc_3D, -3.507),
    "hotrg64z2": (lambda: tk.HOTRG(tk.classical_ising(tk.Z2Irrep)), 64, 7, 4.0, tk.ising_βc, tk.f_onsager),
    "atrg64": (lambda: tk.ATRG(tk.classical_ising(tk.Trivial)), 64, 6, 4.0, tk.ising_βc, tk.f_onsager),
    "hotrg32": (lambda: tk.HOTRG(tk.classical_ising(tk.Trivial)), 32, 6, 4.0, tk.ising_βc, tk.f_onsager),
}
for name in which:
    mk, chi, nsteps, sf, beta, fexact = cases[name]
    s = mk()
    trunc = tk.truncrank(chi)
    data = [s.finalize()]
    times = []
    ctx.reset_counters()
    for it in range(nsteps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.step(trunc); data.append(s.finalize())
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    f = tk.free_energy(data, beta, scalefactor=sf)
    c = ctx.counters()
    extra = {}
    import ctypes as C
    for nm in ("subspace_eigh", "subspace_svd", "subspace_fallbacks", "grouped_gemm_launches"):
        v = C.c_double(); ctx.call("tnr_get_counter", nm.encode(), C.byref(v)); extra[nm] = int(v.value)
    print(json.dumps({"config": name, "chi": chi, "dims": list(s.T.dims), "step_s": [round(t, 3) for t in times],
                      "steady_s_per_step": round(min(times[-2:]), 3), "free_energy": f,
                      "rel_err_vs_exact_after_%d_steps" % nsteps: abs((f - fexact) / fexact),
                      "gemm_tflop_total": round(c["gemm_flops"] / 1e12, 2), "launches": c["launches"], **extra}), flush=True)
    del s
    torch.cuda.empty_cache()
