"""Interface behaviour of the host mirror on the GPU: run! options, stopping criteria, error
behaviour of the C ABI, anisotropic leg dimensions, beta sweeps, large-chi ATRG_3D."""
import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def test_run_without_initial_finalize(tk):
    T = tk.classical_ising(tk.Trivial, 0.4)
    got = tk.run(tk.TRG(T), tk.truncrank(6), tk.maxiter(4), finalize_beginning=False, verbosity=0)
    ref = o.run(o.TRG(T), 6, 4, finalize_beginning=False)
    assert len(got) == 4
    assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= RTOL


def test_convcrit_stops_like_the_reference_loop(tk):
    # example/example.jl: stop when |log(data[end])| 2^-steps <= delta or after maxiter
    f = lambda steps, data: abs(np.log(data[-1]) * 2.0 ** (-steps))
    crit = tk.convcrit(1e-3, f) & tk.maxiter(20)
    got = tk.run(tk.BTRG(tk.classical_ising(1.0)), tk.truncrank(8), crit, verbosity=0)
    s = o.BTRG(o.classical_ising_z2basis(1.0))
    ref, steps = [s.finalize()], 0
    while True:
        s.step(8)
        ref.append(s.finalize())
        steps += 1
        if not (1e-3 < f(steps, ref) and steps < 20):
            break
    assert len(got) == len(ref) and len(got) < 21
    assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= RTOL


def test_anisotropic_legs(tk):
    rng = np.random.default_rng(3)
    T = rng.random((3, 4, 4, 3)) + 0.1
    for name, chi in (("TRG", 5), ("BTRG", 5), ("HOTRG", 4), ("ATRG", 5)):
        got = tk.run(getattr(tk, name)(T), tk.truncrank(chi), tk.maxiter(3), verbosity=0)
        ref = o.run(getattr(o, name)(T), chi, 3)
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= RTOL, name


def test_error_behaviour(tk, ctx):
    import ctypes as C

    from tnrkit.jl_b200 import _lib
    t = tk.DeviceTensor.from_numpy(np.ones((2, 3, 4)))
    with pytest.raises(tk.TNRCudaError):
        t.permute((0, 0, 1))                       # not a permutation
    with pytest.raises(tk.TNRCudaError):
        tk.contract(t, "abc", tk.DeviceTensor.from_numpy(np.ones((5, 2))), "cx", "abx")  # 4 != 5
    with pytest.raises(tk.TNRCudaError):
        tk.svd_trunc(t, 3, 2)                      # ncod must leave a domain
    with pytest.raises(TypeError):
        tk.TRG(np.ones((2, 2, 2)))                 # wrong number of legs
    with pytest.raises(TypeError):
        tk.run(tk.TRG(tk.classical_ising()), 16, tk.maxiter(2))   # not a truncation strategy
    with pytest.raises(tk.TNRCudaError):
        ctx.call("tnr_finalize_2d", t.ptr, _lib.i64((2, 3, 3, 4)), C.byref(C.c_double()))
    assert b"finalize" in ctx.lib.tnr_last_error(ctx.h)
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("no_such_option", 1)


def test_beta_sweep_single_process(tk):
    betas = [0.3, 0.44, 0.6]
    res = tk.beta_sweep(tk.TRG, lambda b: tk.classical_ising(b), betas, tk.truncrank(6), tk.maxiter(4))
    assert len(res) == 3
    for b, data in zip(betas, res):
        ref = o.run(o.TRG(o.classical_ising_z2basis(b)), 6, 4)
        assert np.max(np.abs(np.array(data) - ref) / np.abs(ref)) <= RTOL


def test_atrg3d_chi12_large_matrices(tk):
    """ATRG_3D at chi = 12 (the reference's own test size): 1728 x 1728 SVD operands, top-12 by
    the subspace solver; three RG steps against the oracle."""
    T = tk.classical_ising_3D()
    got = np.array(tk.run(tk.ATRG_3D(T), tk.truncrank(12), tk.maxiter(3), verbosity=0))
    ref = np.array(o.run(o.ATRG_3D(np.asarray(T)), 12, 3))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
