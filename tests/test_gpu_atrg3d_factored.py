"""Device twin of tests/test_atrg3d_factored_emulated.py: the factored ATRG_3D step
(tnrkit.jl_b200/atrg3d_factored.py) through the real C ABI -- `tnr_orth_r`, the implicit products,
the subspace-iteration SVD, chunked TSQR -- against the oracle, against the dense device step
(`tnr_atrg3d_step`) and, on a box with 2 GPUs, sharded over NCCL."""
import json
import os
import socket
import sys

import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_orth_r_gram_identity(tk, ctx):
    """tnr_orth_r: R^T R = T^T T (R is defined up to an orthogonal factor on the left)."""
    from tnrkit.jl_b200 import _lib

    rng = np.random.default_rng(11)
    a = rng.standard_normal((6, 5, 4, 3, 2)) * np.logspace(0, -5, 6).reshape(3, 2)
    T = tk.DeviceTensor.from_numpy(a)
    R = tk.DeviceTensor.empty((6, 6), 1, ctx)
    ctx.call("tnr_orth_r", T.ptr, 5, _lib.i64(a.shape), 3, R.ptr)
    A = a.reshape((120, 6), order="F")
    r = R.to_numpy()
    assert np.abs(r.T @ r - A.T @ A).max() <= 1e-13 * np.abs(A.T @ A).max()
    with pytest.raises(tk.TNRCudaError):
        ctx.call("tnr_orth_r", T.ptr, 5, _lib.i64(a.shape), 1, R.ptr)   # 6 x 120: not tall


@pytest.mark.parametrize("chi,n,block", [(4, 3, None), (6, 3, None), (6, 3, 14), (10, 2, 24)])
def test_factored_atrg3d_matches_oracle_on_device(tk, ctx, chi, n, block):
    from tnrkit.jl_b200 import atrg3d_factored as af

    T = tk.classical_ising_3D(tk.Trivial)
    s = tk.ATRG_3D(T, factored=True, block=block)
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    ref = np.array(o.run(o.ATRG_3D(T), chi, n))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    if block is not None:
        assert all(not st["dense"] for st in af.LAST_STATS["svd"])
    assert s.T.dims == (chi,) * 6


def test_factored_equals_dense_device_step(tk, ctx):
    """chi = 12 (the reference's ATRG_3D testset size, test/schemes.jl): the factored step with
    chunked TSQR against `tnr_atrg3d_step` on the same device, both at FP64."""
    from tnrkit.jl_b200 import atrg3d_factored as af

    T = tk.classical_ising_3D(tk.Trivial)
    chi, n = 12, 3
    dense = np.array(tk.run(tk.ATRG_3D(T, factored=False), tk.truncrank(chi), tk.maxiter(n),
                            verbosity=0))
    s = tk.ATRG_3D(T, factored=True, max_chunk_elems=12 ** 5 * 5)     # chunks of 5, 5, 2
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert af.LAST_STATS["chunks"]["AX"] == [3]
    assert np.max(np.abs(got - dense) / np.abs(dense)) <= RTOL
    # and both against the committed oracle vector (tests/golden/make_golden.py)
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_norms.json")))
    ref = np.array(g["ATRG_3D_ising_trivial_chi12_it6"])[: n + 1]
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL
    assert np.max(np.abs(dense - ref) / np.abs(ref)) <= RTOL


@pytest.mark.parametrize("chi", [16, 24])
def test_full_size_parity_dense_and_factored_device_steps(tk, ctx, chi):
    """Sizes the dense numpy oracle cannot reach in seconds (chi = 24: 1.5 GB tensors): the DENSE
    device step (`tnr_atrg3d_step`) and the factored device step against the vector computed on the
    CPU by the factored step over LAPACK (tests/golden/make_golden_factored.py) -- three
    implementations (dense/Jacobi/GPU, factored/Jacobi/GPU, factored/LAPACK/CPU) of atrg3d.jl."""
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "factored_cpu_norms.json")))
    ref = np.array(g[f"ATRG_3D_ising_trivial_chi{chi}_it3"])
    if chi == 16:
        # chi = 16 is also within reach of the dense numpy ORACLE (380 s of host time): its norm
        # list (tests/golden/baseline_sizes.json, conditioning-checked) is the reference there
        o16 = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_sizes.json")))
        o16 = o16["ATRG_3D_ising_trivial_chi16_it4"]
        assert o16["valid"] and np.max(np.abs(np.array(o16["norms"][:4]) - ref) / np.abs(ref)) <= 1e-12
        ref = np.array(o16["norms"][:4])
    T = tk.classical_ising_3D(tk.Trivial)
    dense = np.array(tk.run(tk.ATRG_3D(T, factored=False), tk.truncrank(chi), tk.maxiter(3),
                            verbosity=0))
    assert np.max(np.abs(dense - ref) / np.abs(ref)) <= RTOL
    for rfactor in ("tsqr", "gram", "gram_eigh"):
        fact = np.array(tk.run(tk.ATRG_3D(T, factored=True, rfactor=rfactor), tk.truncrank(chi),
                               tk.maxiter(3), verbosity=0))
        assert np.max(np.abs(fact - ref) / np.abs(ref)) <= RTOL, rfactor


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, chi, n, rfactor, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import tnrkit.jl_b200 as tk

    from tnrkit.jl_b200 import atrg3d_factored as af

    s = tk.ATRG_3D(tk.classical_ising_3D(tk.Trivial), shard=True, max_chunk_elems=chi ** 5,
                   rfactor=rfactor, block=None if chi < 6 else 2 * chi + 2)
    got = tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    q.put((rank, (got, json.loads(json.dumps(af.LAST_STATS, default=str)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("chi,n,rfactor", [(5, 2, "tsqr"),     # ragged ownership of the open bond: 3 + 2
                                           (10, 3, "gram")])   # sharded products of the subspace iteration (block 22 of 1000), dealt projector pairs
def test_factored_atrg3d_sharded_two_gpus(chi, n, rfactor):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tnrkit.jl_b200 as tk
    from conditioning import checked_oracle_norms

    ref = np.array(checked_oracle_norms("ATRG_3D", chi, n))       # refuses ill-conditioned cases
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, chi, n, rfactor, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(2):
        got = np.array(res[r][0])
        assert np.max(np.abs(got - ref) / np.abs(ref)) <= RTOL, r
        assert res[r][1]["rfactor"] == rfactor
    assert res[0][0] == res[1][0]          # replicas stay bit-identical
    if rfactor == "gram":
        assert res[0][1]["projector_owners"] == [0, 1]
        assert all(not st["dense"] and st["iterations"] >= 2 for st in res[0][1]["svd"])
