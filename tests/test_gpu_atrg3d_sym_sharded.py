"""Device twins of the chunked / sharded block-sparse ATRG_3D step (symmetric.py:
_atrg3d_tail_sharded; reference body: src/schemes/atrg3d.jl:34-83 on the Z2 tensor of
test/schemes.jl:8).  CPU twins on the emulated primitives:
tests/test_host_sequencing_emulated.py::test_emulated_atrg3d_chunked_tail_matches_oracle and
tests/test_multiproc_gloo.py::test_block_sparse_atrg3d_sharded_world2."""
import os
import socket
import sys

import numpy as np
import pytest

import tnr_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# chi = 8 is NOT a parity case: truncrank(8) cuts a degenerate multiplet of the 3D Ising ATRG
# spectrum and the oracle's own norms move by 8e-3 under a 1e-14 perturbation (round 1's red test);
# chi = 4, 6, 10 move by <= 2e-13.  tests/conditioning.py refuses ill-conditioned cases.
@pytest.mark.parametrize("chi,n,chunk", [(4, 3, 1), (6, 2, 2), (10, 3, 3)])
def test_block_sparse_atrg3d_chunked_tail_on_device(tk, chi, n, chunk):
    from conditioning import require_well_conditioned
    from tnrkit.jl_b200 import symmetric

    T = tk.classical_ising_3D()
    ref = require_well_conditioned(lambda t: o.run(o.ATRG_3D(t), chi, n), np.asarray(T),
                                   f"ATRG_3D chi={chi} n={n}")
    s = tk.ATRG_3D(T, symmetric=True, sym_chunk=chunk)
    got = np.array(tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0))
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-10
    plan = symmetric.LAST_PLAN["atrg3d"]
    assert plan["world"] == 1 and plan["chunks_AX"] >= 2 and plan["chunks_YD"] >= 2
    base = np.array(tk.run(tk.ATRG_3D(T, symmetric=True), tk.truncrank(chi), tk.maxiter(n),
                           verbosity=0))
    assert np.max(np.abs(got - base) / np.abs(base)) <= 1e-10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, chi, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    import tnrkit.jl_b200 as tk
    from tnrkit.jl_b200 import symmetric

    s = tk.ATRG_3D(tk.classical_ising_3D(), symmetric=True, shard=True)
    got = tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    q.put((rank, got, dict(symmetric.LAST_PLAN["atrg3d"])))
    dist.barrier()
    dist.destroy_process_group()


def test_block_sparse_atrg3d_sharded_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    chi, n = 6, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, chi, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (got, plan) for r, got, plan in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import tnrkit.jl_b200 as tk

    from conditioning import require_well_conditioned

    ref = require_well_conditioned(lambda t: o.run(o.ATRG_3D(t), chi, n),
                                   np.asarray(tk.classical_ising_3D()), f"ATRG_3D chi={chi} n={n}")
    for r in range(2):
        got, plan = res[r]
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10, r
        assert plan["world"] == 2 and plan["my_chunks"] >= 2
    assert res[0][0] == res[1][0]
