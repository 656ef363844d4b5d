"""TNR schemes behind the unchanged `run!(scheme, truncrank(chi), maxiter(n))` API.

Mirrors /root/reference/src/schemes/{tnrscheme,trg,btrg,hotrg,atrg,hotrg3d,atrg3d}.jl and
src/utility/finalize.jl: same constructor names, same fields (`T`, `S1`, `S2`, `k`), same
`step!` / `finalize!` split, same `run!` loop and return value (the list of norms).  All
arithmetic happens in libtnrcuda through the C ABI; this file only sequences calls."""
from __future__ import annotations

import ctypes as C
import logging
import math
import time

import numpy as np

from . import _lib
from .stopping import maxiter, stopcrit
from .tensor import DeviceTensor
from .truncation import TruncationStrategy, truncrank

log = logging.getLogger("tnrkit.jl_b200")


def _chi(trunc) -> int:
    if not isinstance(trunc, truncrank):
        raise NotImplementedError(
            f"{trunc!r}: only truncrank(chi) is implemented on the B200 path")
    return trunc.chi


class Finalizer:
    """Finalizer(f!, E): function applied after every step (tnrscheme.jl:16-26)."""

    def __init__(self, f, E=float):
        self.f = f
        self.E = E


class TNRScheme:
    kind = None
    nlegs = 4

    def __init__(self, T, ctx=None):
        if isinstance(T, DeviceTensor):
            self.T = T
        else:
            T = np.asarray(T, dtype=np.float64)
            if T.ndim != self.nlegs:
                raise TypeError(f"{type(self).__name__} expects a tensor with {self.nlegs} legs")
            self.T = DeviceTensor.from_numpy(T, 2, ctx)
        self.ctx = self.T.ctx

    # -- step!/finalize! ------------------------------------------------------
    def _out_dims(self, chi):
        n = self.nlegs
        out = (C.c_int64 * n)()
        rc = self.ctx.lib.tnr_step_out_dims(self.kind, _lib.i64(self.T.dims), chi, out)
        self.ctx.check(rc, "tnr_step_out_dims")
        return tuple(out)

    _step_fn = None

    def step(self, trunc):
        chi = _chi(trunc)
        od = self._out_dims(chi)
        out = DeviceTensor.empty(od, 2, self.ctx)
        dims_out = (C.c_int64 * self.nlegs)()
        self.ctx.call(self._step_fn, self.T.ptr, _lib.i64(self.T.dims), chi, out.ptr, dims_out)
        assert tuple(dims_out) == od, (tuple(dims_out), od)
        self.T = out
        return self

    def finalize(self):
        n = C.c_double()
        fn = "tnr_finalize_2d" if self.nlegs == 4 else "tnr_finalize_3d"
        self.ctx.call(fn, self.T.ptr, _lib.i64(self.T.dims), C.byref(n))
        return n.value

    def __repr__(self):
        return f"{type(self).__name__}(T: {self.T.dims})"




def _as_symmetric(T, ctx=None):
    """ChargedArray (host, charge basis) -> block-sparse SymTensor on the device."""
    from .symmetric import Leg, SymTensor

    legs = []
    for ch, sg in zip(T.charges, T.signs):
        if list(ch) != sorted(ch):
            raise ValueError("ChargedArray legs must be ordered by charge")
        sect = {}
        for q in ch:
            sect[q] = sect.get(q, 0) + 1
        legs.append(Leg(sect, sg))
    return SymTensor.from_dense(np.asarray(T), T.N, legs, ctx)


class _SymmetricMixin:
    """TRG / BTRG accept Z_N-symmetric tensors (`classical_ising(Z2Irrep, ...)`,
    `classical_potts(ZNIrrep[q], ...)`) and then run block-sparse: per-sector SVD with
    sector-global truncrank, grouped per-sector GEMM.  `symmetric=False` forces the dense path."""

    def _init_sym(self, T, symmetric, ctx):
        from .symmetric import SymTensor

        if isinstance(T, SymTensor):
            self.T, self.ctx, self.sym = T, T.ctx, True
            return True
        if symmetric is not False and getattr(T, "charges", None) is not None:
            self.T = _as_symmetric(T, ctx)
            self.ctx, self.sym = self.T.ctx, True
            return True
        self.sym = False
        return False


class TRG(_SymmetricMixin, TNRScheme):
    """Tensor Renormalization Group (trg.jl)."""
    kind = _lib.TNR_TRG
    _step_fn = "tnr_trg_step"

    def __init__(self, T, ctx=None, symmetric=None):
        if not self._init_sym(T, symmetric, ctx):
            TNRScheme.__init__(self, T, ctx)

    def step(self, trunc):
        if not self.sym:
            return TNRScheme.step(self, trunc)
        from .symmetric import trg_step_sym

        self.T = trg_step_sym(self.T, _chi(trunc))
        return self

    def finalize(self):
        if not self.sym:
            return TNRScheme.finalize(self)
        from .symmetric import sym_trace_2d

        n = abs(sym_trace_2d(self.T))
        self.T.scale(1.0 / n)
        return n


class _Sym2D(_SymmetricMixin, TNRScheme):
    """2D schemes whose step! also exists on block-sparse Z_N tensors."""
    _sym_step = None

    def __init__(self, T, ctx=None, symmetric=None):
        if not self._init_sym(T, symmetric, ctx):
            TNRScheme.__init__(self, T, ctx)

    def step(self, trunc):
        if not self.sym:
            return TNRScheme.step(self, trunc)
        from . import symmetric

        self.T = getattr(symmetric, self._sym_step)(self.T, _chi(trunc))
        return self

    def finalize(self):
        if not self.sym:
            return TNRScheme.finalize(self)
        from .symmetric import sym_trace_2d

        n = abs(sym_trace_2d(self.T))
        self.T.scale(1.0 / n)
        return n


class HOTRG(_Sym2D):
    """Higher-Order TRG (hotrg.jl)."""
    kind = _lib.TNR_HOTRG
    _step_fn = "tnr_hotrg_step"
    _sym_step = "hotrg_step_sym"


class ATRG(_Sym2D):
    """Anisotropic TRG (atrg.jl)."""
    kind = _lib.TNR_ATRG
    _step_fn = "tnr_atrg_step"
    _sym_step = "atrg_step_sym"


class _Sym3D(_SymmetricMixin):
    """3D schemes on block-sparse Z_N tensors.  Opt-in (`symmetric=True`, or a SymTensor) in this
    round: without it a charged model tensor is stored densely in the charge basis, which gives
    the same numbers unless the cut splits an exactly degenerate multiplet."""
    _sym_step = None

    def _init_sym3d(self, T, symmetric, ctx):
        from .symmetric import SymTensor

        if isinstance(T, SymTensor) or symmetric:
            if not isinstance(T, SymTensor) and getattr(T, "charges", None) is None:
                raise TypeError("symmetric=True needs a tensor in a charge basis "
                                "(classical_ising_3D(Z2Irrep, ...))")
            return self._init_sym(T, True, ctx)
        self.sym = False
        return False

    def _sym_finalize(self):
        from .symmetric import sym_trace_3d

        n = abs(sym_trace_3d(self.T))
        self.T.scale(1.0 / n)
        return n


class ATRG_3D(_Sym3D, TNRScheme):
    """3D Anisotropic TRG (atrg3d.jl).

    `factored=True` keeps the tensor in the two-factor form `G * H` in which `_step!` produces it
    and never builds a chi^6 object (tnrkit.jl_b200/atrg3d_factored.py): truncated SVDs by
    subspace iteration on the implicit operator, R factors by TSQR over chunks of an open bond
    (`rfactor="tsqr"`, default below chi = 40) or by pivoted Cholesky of the Gram matrices of the
    matricizations (`rfactor="gram"`, default from chi = 40).
    `factored=None` switches to it by itself when the dense step's chi^6 tensors would not fit in
    HBM (chi >= 40).  With `shard=True` under torch.distributed (one process per GPU) the chunks
    of the open bond are divided between the ranks (all-gather of the chunk R factors and of
    H / G; nothing else is exchanged)."""
    kind = _lib.TNR_ATRG_3D
    nlegs = 6
    _step_fn = "tnr_atrg3d_step"
    DENSE_BYTES_LIMIT = 150e9

    def __init__(self, T, ctx=None, symmetric=None, factored=None, shard=None, group=None,
                 max_chunk_elems=1 << 30, tol=1e-13, block=None, rfactor=None, sym_chunk=None):
        self._F = None
        self.block = block
        self.rfactor = rfactor
        self.factored = factored
        self.group = group
        self.max_chunk_elems = int(max_chunk_elems)
        self.tol = float(tol)
        if self._init_sym3d(T, symmetric, ctx):
            if factored:
                raise NotImplementedError("factored ATRG_3D on block-sparse tensors")
            # block-sparse: shard=True deals the chunks of the open bond of AX / YD to the ranks
            # (symmetric.py: _atrg3d_tail_sharded); `sym_chunk` chunks it on one process as well
            self.factored = False
            self.shard = bool(shard)
            self.sym_chunk = sym_chunk
            return
        TNRScheme.__init__(self, T, ctx)
        if shard and factored is False:
            raise ValueError("ATRG_3D: shard=True needs the factored step")
        if shard:
            self.factored = True
        self.shard = bool(shard)
        if self.factored:
            self._to_factored()

    @classmethod
    def wants_factored(cls, chi: int) -> bool:
        """The dense step holds up to 6 tensors of chi^6 doubles at once (T, AX, YD, a permuted
        copy, the work copy of the R factorization, the result); beyond DENSE_BYTES_LIMIT (of
        the 180 GB of HBM) the factored step takes over: chi >= 40."""
        return 6 * 8.0 * float(chi) ** 6 > cls.DENSE_BYTES_LIMIT

    # `T` stays the reference's field: with the factored state it is materialised on request
    @property
    def T(self):
        if self._F is not None:
            return self._F.to_dense()
        return self._T

    @T.setter
    def T(self, value):
        self._T = value
        self._F = None

    @property
    def factors(self):
        """The TwoFactor state of the factored step (None on the dense path)."""
        return self._F

    def _to_factored(self):
        if self._F is None:
            from .atrg3d_factored import TwoFactor

            self._F = TwoFactor.from_dense(self._T)
            self._T = None

    def step(self, trunc):
        chi = _chi(trunc)
        if self.sym:
            from .symmetric import atrg3d_step_sym

            shard = None
            if self.shard:
                import torch.distributed as dist

                shard = (dist.get_rank(self.group), dist.get_world_size(self.group), self.group)
            self.T = atrg3d_step_sym(self.T, chi, shard, self.sym_chunk)
            return self
        if self.factored is None:
            self.factored = self.wants_factored(chi)
        if not self.factored:
            return TNRScheme.step(self, trunc)
        from .atrg3d_factored import atrg3d_step_factored

        self._to_factored()
        # R factors: chunked Householder TSQR (the reference's left_orth, literally) while it is
        # affordable; from the chi where the dense step no longer fits (chi >= 40) its 4 x 2 chi^8
        # flop and chi^2-column panels over chi^4-row chunks cost minutes per step (chi = 48:
        # > 100 s per `_step!`), and the factors come from the Gram matrices instead (O(chi^7))
        rf = self.rfactor or ("gram" if self.wants_factored(chi) else "tsqr")
        self._F = atrg3d_step_factored(self._F, chi, max_chunk_elems=self.max_chunk_elems,
                                       shard=self.shard, group=self.group, tol=self.tol,
                                       block=self.block, rfactor=rf)
        return self

    def finalize(self):
        if self.sym:
            return self._sym_finalize()
        if self._F is not None:
            n = abs(self._F.trace_3d())
            self._F.scale(1.0 / n)
            return n
        return TNRScheme.finalize(self)

    def __repr__(self):
        if self._F is not None:
            return f"ATRG_3D(T: {self._F.dims}, factored, bond {self._F.bond_dim})"
        return TNRScheme.__repr__(self)


class BTRG(_SymmetricMixin, TNRScheme):
    """Bond-weighted TRG (btrg.jl): fields T, S1 (vertical bonds), S2 (horizontal), k."""
    kind = _lib.TNR_BTRG

    def __init__(self, T, k=-0.5, ctx=None, symmetric=None):
        self.k = float(k)
        if self._init_sym(T, symmetric, ctx):
            from .symmetric import identity_weights

            self.S1 = identity_weights(self.T.legs[1], self.ctx)
            self.S2 = identity_weights(self.T.legs[0], self.ctx)
            return
        TNRScheme.__init__(self, T, ctx)
        # S1 = id(space(T,2)), S2 = id(space(T,1)); stored as diagonals
        self.S1 = DeviceTensor.from_numpy(np.ones(self.T.dims[1]), 1, self.ctx)
        self.S2 = DeviceTensor.from_numpy(np.ones(self.T.dims[0]), 1, self.ctx)
        self.k = float(k)

    def step(self, trunc):
        chi = _chi(trunc)
        if self.sym:
            from .symmetric import btrg_step_sym

            self.T, self.S1, self.S2 = btrg_step_sym(self.T, self.S1, self.S2, self.k, chi)
            return self
        od = self._out_dims(chi)
        out = DeviceTensor.empty(od, 2, self.ctx)
        s1 = DeviceTensor.empty((od[1],), 1, self.ctx)
        s2 = DeviceTensor.empty((od[0],), 1, self.ctx)
        dims_out = (C.c_int64 * 4)()
        self.ctx.call("tnr_btrg_step", self.T.ptr, _lib.i64(self.T.dims), self.S1.ptr,
                      self.S2.ptr, self.k, chi, out.ptr, dims_out, s1.ptr, s2.ptr)
        assert tuple(dims_out) == od, (tuple(dims_out), od)
        self.T, self.S1, self.S2 = out, s1, s2
        return self

    def finalize(self):
        if self.sym:
            from .symmetric import sym_trace_2d

            n = abs(sym_trace_2d(self.T, self.S2, self.S1))
            self.T.scale(1.0 / n)
            return n
        n = C.c_double()
        self.ctx.call("tnr_finalize_btrg", self.T.ptr, _lib.i64(self.T.dims), self.S1.ptr,
                      self.S2.ptr, C.byref(n))
        return n.value


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of the open bond owned by `rank`: equal blocks when world | n, otherwise
    the first n mod world ranks own one index more (blocks differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_last_leg(buf, dims, group=None):
    """All-gather of a column-major tensor sharded along its LAST leg (contiguous slabs).
    Every rank holds the full-size buffer with only its own slab valid."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = int(dims[-1])
    slab = math.prod(int(d) for d in dims[:-1])
    if n % world == 0:
        per = n // world
        mine = buf[rank * per * slab:(rank + 1) * per * slab]
        dist.all_gather_into_tensor(buf[: n * slab], mine, group=group)
    else:
        for r in range(world):
            lo, hi = shard_range(n, r, world)
            if hi > lo:
                dist.broadcast(buf[lo * slab: hi * slab], src=dist.get_global_rank(group, r)
                               if group is not None else r, group=group)
    return buf


class HOTRG_3D(_Sym3D, TNRScheme):
    """3D Higher-Order TRG (hotrg3d.jl).

    With `torch.distributed` initialised (one process per GPU) and `shard=True`, every
    z-compression shards the chi^11 contraction along the new open x-bond: rank r computes
    T'[..., f] for its block of f and the blocks are all-gathered over NCCL (no other
    data-path collective)."""
    kind = _lib.TNR_HOTRG_3D
    nlegs = 6
    _step_fn = "tnr_hotrg3d_step"
    PERM = (5, 3, 1, 2, 0, 4)  # ((6,4),(2,3,1,5))

    def __init__(self, T, ctx=None, shard=None, group=None, peer_scatter=None, symmetric=None,
                 max_chunk_elems=1 << 29, split_projectors=True):
        self.group = group
        # sharded runs: deal the four projector eigendecompositions to the ranks instead of
        # repeating them on every rank (False = round-1 behaviour, for A/B measurements)
        self.split_projectors = split_projectors if split_projectors == "force" else bool(split_projectors)
        self.max_chunk_elems = int(max_chunk_elems)
        is_sym = self._init_sym3d(T, symmetric, ctx)
        if not is_sym:
            TNRScheme.__init__(self, T, ctx)
        if shard is None:
            try:
                import torch.distributed as dist

                shard = dist.is_available() and dist.is_initialized() and \
                    dist.get_world_size(group) > 1
            except Exception:
                shard = False
        self.shard = bool(shard)
        if is_sym:
            return
        # peer_scatter: publish every T' slab to all ranks with NVLink stores from the producing
        # kernel (torch symmetric memory provides the peer-mapped buffers) instead of an NCCL
        # all-gather afterwards.  None = try, fall back to NCCL when symmetric memory is absent.
        self.peer_scatter = peer_scatter
        self._symm = None      # (buffers, handles) of the two ping-pong symmetric T' buffers
        self._symm_turn = 0

    def _symm_buffers(self, nelem):
        """Two symmetric-memory buffers of at least nelem doubles (allocated once, collectively)."""
        import torch
        import torch.distributed as dist

        if self._symm is not None and self._symm[0][0].numel() >= nelem:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem

        grp = self.group if self.group is not None else dist.group.WORLD
        bufs, hdls = [], []
        for _ in range(2):
            t = symm_mem.empty(nelem, dtype=torch.float64, device=self.ctx.torch_device)
            hdls.append(symm_mem.rendezvous(t, group=grp))
            bufs.append(t)
        self._symm = (bufs, hdls)
        return self._symm

    def _projector_halves(self, chi, od, world, rank):
        """The four truncated eigendecompositions of a z-compression (hotrg3d.jl:83-108), dealt
        round-robin to the ranks: half h is computed by rank h % world
        (`tnr_hotrg3d_proj_half`) and broadcast -- 4 small messages (D^2 x chi doubles) instead
        of every rank repeating the Gram contractions and eigensolves.  Every half is computed
        by ONE rank with the kernels a single-GPU run uses, so the result is bit-identical."""
        import torch
        import torch.distributed as dist

        d = self.T.dims
        sx = d[5] * d[5] * od[3] + 1
        sy = d[4] * d[4] * od[2] + 1
        sizes = (sx, sx, sy, sy)
        offs = (0, sx, 2 * sx, 2 * sx + sy)
        buf = torch.empty(2 * sx + 2 * sy, dtype=torch.float64, device=self.ctx.torch_device)
        for h in range(4):
            if h % world == rank:
                self.ctx.call("tnr_hotrg3d_proj_half", self.T.ptr, _lib.i64(d), chi, h,
                              C.c_void_p(buf.data_ptr() + 8 * offs[h]))
        if world > 1:
            self.ctx.synchronize()      # engine stream -> torch stream hand-over
            for h in range(4):
                src = h % world
                if self.group is not None:
                    src = dist.get_global_rank(self.group, src)
                dist.broadcast(buf[offs[h]: offs[h] + sizes[h]], src=src, group=self.group)
            torch.cuda.current_stream().synchronize()
        return buf

    def _substep(self, chi):
        import torch.distributed as dist

        d = self.T.dims
        od = (d[0], d[1], min(chi, d[4] * d[4]), min(chi, d[5] * d[5]),
              min(chi, d[4] * d[4]), min(chi, d[5] * d[5]))
        if self.shard:
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        else:
            world, rank = 1, 0
        lo, hi = shard_range(od[5], rank, world)
        dims_out = (C.c_int64 * 6)()
        nelem = math.prod(od)
        use_peers = world > 1 and self.peer_scatter is not False
        if use_peers:
            try:
                # size the symmetric buffers for the saturated tensor so they are allocated once
                bufs, hdls = self._symm_buffers(max(nelem, chi ** 6))
            except Exception as e:  # symmetric memory unavailable on this platform
                if self.peer_scatter:
                    raise
                log.warning("symmetric memory unavailable (%s); using the NCCL all-gather", e)
                self.peer_scatter = False
                use_peers = False
        split = self.split_projectors == "force" or (world > 1 and self.split_projectors)
        halves = self._projector_halves(chi, od, world, rank) if split else None
        if use_peers:
            turn = self._symm_turn
            self._symm_turn ^= 1
            buf, hdl = bufs[turn], hdls[turn]
            ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            if split:
                self.ctx.call("tnr_hotrg3d_contract", self.T.ptr, _lib.i64(d), chi,
                              C.c_void_p(halves.data_ptr()), ptrs, world, rank, dims_out, lo, hi)
            else:
                self.ctx.call("tnr_hotrg3d_substep_peers", self.T.ptr, _lib.i64(d), chi, ptrs,
                              world, rank, dims_out, lo, hi)
            hdl.barrier()  # all ranks' peer stores have landed in this buffer
            out = DeviceTensor(buf, od, 2, self.ctx)
            self.T = out.permute(self.PERM)
            return
        out = DeviceTensor.empty(od, 2, self.ctx)
        if split:
            ptrs = (C.c_void_p * 1)(int(out.buf.data_ptr()))
            self.ctx.call("tnr_hotrg3d_contract", self.T.ptr, _lib.i64(d), chi,
                          C.c_void_p(halves.data_ptr()), ptrs, 1, 0, dims_out, lo, hi)
        else:
            self.ctx.call("tnr_hotrg3d_substep", self.T.ptr, _lib.i64(d), chi, out.ptr, dims_out,
                          lo, hi)
        if world > 1:
            allgather_last_leg(out.buf, od, self.group)
        self.T = out.permute(self.PERM)

    def step(self, trunc):
        chi = _chi(trunc)
        if self.sym:
            from .symmetric import hotrg3d_step_sym

            sh = None
            if self.shard:
                import torch.distributed as dist

                sh = (dist.get_rank(self.group), dist.get_world_size(self.group), self.group)
            self.T = hotrg3d_step_sym(self.T, chi, self.max_chunk_elems, sh)
            return self
        for _ in range(3):
            self._substep(chi)
        return self

    def finalize(self):
        return self._sym_finalize() if self.sym else TNRScheme.finalize(self)


def finalize(scheme):
    """finalize!(scheme): normalise the tensor by its trace norm and return the norm."""
    return scheme.finalize()


default_Finalizer = Finalizer(finalize, float)


def run(scheme, trscheme, criterion=None, finalizer=None, finalize_beginning=True, verbosity=1):
    """run!(scheme, trscheme, criterion[, finalizer]; finalize_beginning=true, verbosity=1)
    -- tnrscheme.jl:31-59.  Returns the list of finalizer outputs (norms)."""
    if not isinstance(trscheme, TruncationStrategy):
        raise TypeError("run!: second argument must be a truncation strategy")
    criterion = criterion if criterion is not None else maxiter(100)
    if not isinstance(criterion, stopcrit):
        raise TypeError("run!: third argument must be a stopping criterion")
    finalizer = finalizer or default_Finalizer
    data = []
    if verbosity >= 1:
        log.info("Starting simulation\n %r\n", scheme)
    if finalize_beginning:
        data.append(finalizer.f(scheme))
    steps = 0
    crit = True
    t0 = time.perf_counter()
    while crit:
        if verbosity >= 2:
            log.info("Step %d, data[end]: %s", steps + 1, data[-1] if data else "empty")
        scheme.step(trscheme)
        data.append(finalizer.f(scheme))
        steps += 1
        crit = criterion(steps, data)
    if verbosity >= 1:
        log.info("Simulation finished\n %s\n Elapsed time: %.3fs\n Iterations: %d",
                 criterion.info(steps, data), time.perf_counter() - t0, steps)
    return data


run_ = run  # `run!`


def beta_sweep(scheme_cls, model, betas, trscheme, criterion, **scheme_kwargs):
    """beta-sweep: one independent scheme per GPU, no collectives on the data path.
    Rank r handles betas[r::world]; results are gathered as Python objects at the end."""
    try:
        import torch.distributed as dist

        on = dist.is_available() and dist.is_initialized()
    except Exception:
        on = False
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    mine = {}
    for i in range(rank, len(betas), world):
        kw = dict(scheme_kwargs)
        if scheme_cls in (HOTRG_3D, ATRG_3D):
            kw["shard"] = False
        scheme = scheme_cls(model(betas[i]), **kw)
        mine[i] = run(scheme, trscheme, criterion, verbosity=0)
    if not on or world == 1:
        return [mine[i] for i in range(len(betas))]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    merged = {}
    for g in gathered:
        merged.update(g)
    return [merged[i] for i in range(len(betas))]


def finalize_two_by_two(scheme):
    """finalize_two_by_two!(scheme::Union{TRG,ATRG,HOTRG}) -- src/utility/finalize.jl:17-25:
    n = |T[7 1;5 4] T[4 2;6 7] T[3 6;2 8] T[8 5;1 3]|, T /= n^(1/4), returns n^(1/4).
    BTRG (finalize.jl:27-42): the same ring of four tensors with the diagonal bond weights S2 on
    the bonds 8-2 and 4-7 and S1 on the bonds 3-6 and 5-1; they are absorbed into the legs of
    three of the four tensors (`tnr_axis_scale`).  Block-sparse schemes run the contractions
    sector by sector; only a chi x chi matrix is traced on the host."""
    from .tensor import contract as _contract

    sym = getattr(scheme, "sym", False)
    weighted = hasattr(scheme, "S1")
    if sym:
        from .symmetric import sym_clone, sym_contract

        T = scheme.T
        L = T.legs
        if not all(L[i].sign == -L[j].sign and L[i].same_space(L[j]) for i, j in ((0, 3), (1, 2))):
            raise ValueError("finalize_two_by_two: legs 1/4 and 2/3 must be dual to each other")

        def w(legs):
            if not weighted or not legs:
                return T
            W = sym_clone(T)
            for ax in legs:
                W.scale_leg(ax, scheme.S2 if ax == 0 else scheme.S1)
            return W

        X = sym_contract(w((1,)), "gaed", w((0, 1)), "dbfg", "aebf")
        Y = sym_contract(w((0,)), "cfbh", T, "heac", "fbea")
        n = abs(float(np.trace(sym_contract(X, "aebf", Y, "fbxa", "ex").to_dense())))
    else:
        T = scheme.T
        d = T.dims

        def w(legs):
            if not weighted or not legs:
                return T
            W = T.clone()
            for ax in legs:
                S = scheme.S2 if ax == 0 else scheme.S1
                W.ctx.call("tnr_axis_scale", W.ptr, math.prod(d[:ax]), d[ax],
                           math.prod(d[ax + 1:]), S.ptr, 0, 0.0)
            return W

        # labels: 7=g 1=a 5=e 4=d 2=b 6=f 3=c 8=h
        X = _contract(w((1,)), "gaed", w((0, 1)), "dbfg", "aebf")      # sum over d, g
        Y = _contract(w((0,)), "cfbh", T, "heac", "fbea")              # sum over c, h
        n = abs(float(_contract(X, "aebf", Y, "fbea", "").to_numpy().reshape(-1)[0]))
    f = n ** 0.25
    if sym:
        scheme.T.scale(1.0 / f)
    else:
        scheme.ctx.call("tnr_scale", scheme.T.ptr, scheme.T.size, 1.0 / f)
    return f


two_by_two_Finalizer = Finalizer(finalize_two_by_two, float)
