"""world_size-2 gloo test (CPU) of the N>1 host logic: open-bond slab ownership and the
all-gather along the last leg used by the sharded HOTRG_3D step."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dims, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tnrkit.jl_b200 as tk

    n = dims[-1]
    slab = int(np.prod(dims[:-1]))
    full = torch.arange(slab * n, dtype=torch.float64) * 0.5 + 1.0  # the "true" tensor
    buf = torch.full((slab * n,), float("nan"), dtype=torch.float64)
    lo, hi = tk.shard_range(n, rank, world)
    buf[lo * slab: hi * slab] = full[lo * slab: hi * slab]  # what tnr_hotrg3d_substep fills
    tk.allgather_last_leg(buf, dims)
    ok = bool(torch.equal(buf, full))
    # beta sweep bookkeeping: rank r owns betas[r::world]
    mine = list(range(rank, 5, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, ok, sorted(sum(gathered, []))))
    dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(2, 3, 4), (3, 2, 5)])  # even split and ragged split
def test_allgather_last_leg_world2(dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, betas in res:
        assert ok, f"rank {rank}: gathered tensor differs"
        assert betas == [0, 1, 2, 3, 4]


def _sym_worker(rank, world, port, chi, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tnrkit.jl_b200 as tk
    from abi_emulator import EmulatedContext
    from tnrkit.jl_b200 import _lib, symmetric

    _lib._default_ctx = EmulatedContext()   # numpy emulation of the C-ABI primitives (tests only)
    s = tk.HOTRG_3D(tk.classical_ising_3D(), symmetric=True)
    assert s.sym and s.shard                # shard auto-detected from the process group
    got = tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    q.put((rank, got, dict(symmetric.LAST_PLAN["hotrg3d"])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("chi,n", [(4, 3), (6, 2)])
def test_block_sparse_hotrg3d_sharded_world2(chi, n):
    """Block-sparse (Z2) HOTRG_3D with the F chunks of the open x-bond dealt to two ranks and one
    all-reduce of the flat block buffer per z-compression: norm list == oracle on both ranks."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    import tnrkit.jl_b200 as tk

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sym_worker, args=(r, 2, port, chi, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (got, plan) for r, got, plan in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.array(o.run(o.HOTRG_3D(np.asarray(tk.classical_ising_3D())), chi, n))
    for r in range(2):
        got, plan = res[r]
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10, r
        assert plan["world"] == 2 and plan["chunks"] >= 2 and plan["my_F_chunks"] >= 1
    assert res[0][0] == res[1][0]           # replicas stay bit-identical


def _sym_atrg_worker(rank, world, port, chi, n, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tnrkit.jl_b200 as tk
    from abi_emulator import EmulatedContext
    from tnrkit.jl_b200 import _lib, symmetric

    _lib._default_ctx = EmulatedContext()   # numpy emulation of the C-ABI primitives (tests only)
    s = tk.ATRG_3D(tk.classical_ising_3D(), symmetric=True, shard=True)
    assert s.sym and s.shard
    got = tk.run(s, tk.truncrank(chi), tk.maxiter(n), verbosity=0)
    q.put((rank, got, dict(symmetric.LAST_PLAN["atrg3d"])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("chi,n", [(4, 3), (6, 2)])
def test_block_sparse_atrg3d_sharded_world2(chi, n):
    """Block-sparse (Z2) ATRG_3D with the chunks of the open bond of AX / YD dealt to two ranks:
    TSQR stacks of the chunk R factors and H / G each replicated by one all-reduce of a flat
    block buffer.  Norm list == oracle on both ranks, replicas bit-identical."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tnr_oracle as o

    import tnrkit.jl_b200 as tk

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sym_atrg_worker, args=(r, 2, port, chi, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (got, plan) for r, got, plan in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.array(o.run(o.ATRG_3D(np.asarray(tk.classical_ising_3D())), chi, n))
    for r in range(2):
        got, plan = res[r]
        assert np.max(np.abs(np.array(got) - ref) / np.abs(ref)) <= 1e-10, r
        assert plan["world"] == 2 and plan["chunks_AX"] >= 2 and plan["my_chunks"] >= 2
    assert res[0][0] == res[1][0]           # replicas stay bit-identical
