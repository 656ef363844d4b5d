"""2-GPU NCCL test of the sharded HOTRG_3D step (skipped on a 1-GPU box): the norm list of
the sharded run must equal the single-GPU run and the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, chi, nsteps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import tnrkit.jl_b200 as tk

    import ctypes as C

    ctx = tk.default_context()
    out = {}
    for mode in ("peers", "nccl"):
        s = tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial),
                        peer_scatter=None if mode == "peers" else False)
        assert s.shard
        v0 = C.c_double()
        ctx.call("tnr_get_counter", b"peer_scatter_launches", C.byref(v0))
        out[mode] = tk.run(s, tk.truncrank(chi), tk.maxiter(nsteps), verbosity=0)
        v1 = C.c_double()
        ctx.call("tnr_get_counter", b"peer_scatter_launches", C.byref(v1))
        out[mode + "_peer_launches"] = v1.value - v0.value
    # round-1 behaviour (every rank repeats the projector work) and the single-GPU run
    s = tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), split_projectors=False)
    out["replicated_projectors"] = tk.run(s, tk.truncrank(chi), tk.maxiter(nsteps), verbosity=0)
    s = tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial), shard=False)
    out["single"] = tk.run(s, tk.truncrank(chi), tk.maxiter(nsteps), verbosity=0)
    # the opt-in INT8 engine inside the sharded step (chi = 8: 512^3 chunk contractions)
    ref8 = tk.run(tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial)), tk.truncrank(8), tk.maxiter(2),
                  verbosity=0)
    ctx.set_option("ozaki", 8)
    try:
        g0 = C.c_double()
        ctx.call("tnr_get_counter", b"ozaki_gemms", C.byref(g0))
        oz8 = tk.run(tk.HOTRG_3D(tk.classical_ising_3D(tk.Trivial)), tk.truncrank(8),
                     tk.maxiter(2), verbosity=0)
        g1 = C.c_double()
        ctx.call("tnr_get_counter", b"ozaki_gemms", C.byref(g1))
    finally:
        ctx.set_option("ozaki", 0)
    out["ozaki_used"] = g1.value - g0.value
    out["ozaki_maxrel"] = max(abs(a - b) / abs(b) for a, b in zip(oz8, ref8))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_hotrg3d_sharded_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    chi, nsteps = 5, 3  # chi = 5: ragged split of the open bond over 2 ranks (3 + 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, chi, nsteps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = np.array(o.run(o.HOTRG_3D(o.classical_ising_3D()), chi, nsteps))
    for r in range(2):
        for mode in ("peers", "nccl"):
            got = np.array(res[r][mode])
            assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-10, (r, mode)
        assert res[r]["nccl_peer_launches"] == 0
        assert res[r]["ozaki_used"] > 0 and res[r]["ozaki_maxrel"] <= 1e-11
        # the two exchange mechanisms move the same numbers: bit-identical norm lists
        assert res[r]["peers"] == res[r]["nccl"]
        # projector halves computed on different ranks and broadcast == computed everywhere
        # == the single-GPU run, bit for bit
        assert res[r]["peers"] == res[r]["replicated_projectors"] == res[r]["single"]
    assert res[0]["nccl"] == res[1]["nccl"]  # replicas stay bit-identical
    print("peer-scatter launches per rank:", res[0]["peers_peer_launches"],
          "(0 means symmetric memory was unavailable and the NCCL path was used)")


BETAS_3D = [0.20, 0.2216544, 0.25]


def _sweep_worker(rank, world, port, betas, chi, nsteps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import tnrkit.jl_b200 as tk

    ctx = tk.default_context()
    ctx.reset_counters()
    res2d = tk.beta_sweep(tk.TRG, lambda b: tk.classical_ising(tk.Trivial, b), betas,
                          tk.truncrank(chi), tk.maxiter(nsteps))
    n2d = ctx.counters()["launches"]
    # (3D Ising at beta = 0.40 is deep in the ordered phase: truncrank(4) cuts a degenerate
    #  multiplet there and the oracle itself moves by 2e-5 under a 1e-14 perturbation)
    res3d = tk.beta_sweep(tk.HOTRG_3D, lambda b: tk.classical_ising_3D(tk.Trivial, b), BETAS_3D,
                          tk.truncrank(4), tk.maxiter(2))
    q.put((rank, res2d, res3d, n2d))
    dist.barrier()
    dist.destroy_process_group()


def test_beta_sweep_two_gpus():
    """beta-sweep on 2 processes / 2 GPUs (north star: one independent scheme per GPU, no
    collective on the data path): rank r runs betas[r::2], the norm lists are exchanged as host
    objects at the end, every rank holds all of them and each equals the oracle at its beta."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    betas = [0.40, 0.42, 0.44068679350977147, 0.46, 0.48]
    chi, nsteps = 8, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, betas, chi, nsteps, q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (a, b, n) for r, a, b, n in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]     # everybody holds every result
    assert res[0][2] > 0 and res[1][2] > 0                       # both GPUs did device work
    assert res[0][2] > res[1][2]                                 # rank 0 ran 3 of the 5 betas
    for b, got in zip(betas, res[0][0]):
        ref = np.array(o.run(o.TRG(o.classical_ising(b)), chi, nsteps))
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10, b
    from conditioning import require_well_conditioned

    for b, got in zip(BETAS_3D, res[0][1]):
        ref = require_well_conditioned(lambda t: o.run(o.HOTRG_3D(t), 4, 2),
                                       o.classical_ising_3D(b), f"HOTRG_3D beta={b}")
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10, b
