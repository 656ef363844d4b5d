"""Device-resident tensor storage (the counterpart of the new TensorMap device-storage
backend the north star asks for).  Column major, FP64, data lives in a torch CUDA buffer
(torch is only the allocator / stream / NCCL plumbing); all arithmetic goes through
libtnrcuda."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib


def _torch():
    import torch

    return torch


class DeviceTensor:
    """Dense `Trivial`-sector TensorMap data: legs (codomain..., domain...), first leg fastest."""

    def __init__(self, buf, dims, ncod=None, ctx=None):
        self.buf = buf  # 1-D torch.float64 CUDA tensor with prod(dims) elements (or more)
        self.dims = tuple(int(d) for d in dims)
        self.ncod = ncod
        self.ctx = ctx or _lib.default_context()

    # -- construction -----------------------------------------------------
    @classmethod
    def empty(cls, dims, ncod=None, ctx=None):
        torch = _torch()
        ctx = ctx or _lib.default_context()
        n = max(1, math.prod(int(d) for d in dims))
        buf = torch.empty(n, dtype=torch.float64, device=ctx.torch_device)
        return cls(buf, dims, ncod, ctx)

    @classmethod
    def from_numpy(cls, arr, ncod=None, ctx=None, pinned=None):
        """Uploads a host array indexed arr[leg1, leg2, ...] (any memory order)."""
        torch = _torch()
        a = np.asarray(arr, dtype=np.float64)
        flat = np.ascontiguousarray(np.transpose(a).reshape(-1))  # column-major flattening
        t = cls.empty(a.shape, ncod, ctx)
        src = torch.from_numpy(flat)
        if pinned is not None:
            pinned[: flat.size].copy_(src)
            src = pinned[: flat.size]
        t.buf[: flat.size].copy_(src, non_blocking=pinned is not None)
        return t

    def to_numpy(self):
        n = self.size
        flat = self.buf[:n].cpu().numpy()
        return np.transpose(flat.reshape(tuple(reversed(self.dims))))

    # -- helpers ------------------------------------------------------------
    @property
    def size(self):
        return math.prod(self.dims)

    @property
    def ptr(self):
        return C.c_void_p(self.buf.data_ptr())

    def clone(self):
        return DeviceTensor(self.buf.clone(), self.dims, self.ncod, self.ctx)

    def permute(self, perm):
        """TensorKit.permute: new leg k is old leg perm[k] (0-based here)."""
        out = DeviceTensor.empty([self.dims[p] for p in perm], self.ncod, self.ctx)
        self.ctx.call("tnr_permute", self.ptr, out.ptr, len(self.dims), _lib.i64(self.dims),
                      _lib.i32(perm))
        return out

    def __repr__(self):
        return f"DeviceTensor(dims={self.dims}, device={self.ctx.torch_device})"


def contract(A: DeviceTensor, la: str, B: DeviceTensor, lb: str, lc: str,
             out: DeviceTensor | None = None) -> DeviceTensor:
    """One binary `@tensor` contraction by leg labels on the device (tnr_contract).
    `out`: write into this tensor (e.g. a slab view of a larger buffer) instead of a new one."""
    da = dict(zip(la, A.dims))
    da.update(zip(lb, B.dims))
    od = tuple(da[c] for c in lc)
    if out is None:
        out = DeviceTensor.empty(od, None, A.ctx)
    elif tuple(out.dims) != od:
        raise ValueError(f"contract: out has dims {out.dims}, expected {od}")
    A.ctx.call("tnr_contract", A.ptr, len(A.dims), _lib.i64(A.dims), la.encode(), B.ptr,
               len(B.dims), _lib.i64(B.dims), lb.encode(), out.ptr, lc.encode())
    return out


def svd_trunc(T: DeviceTensor, ncod: int, chi: int):
    """svd_trunc(T; trunc=truncrank(chi)): returns U, S, Vt (DeviceTensors) and eps."""
    m = math.prod(T.dims[:ncod])
    n = math.prod(T.dims[ncod:])
    k = min(chi, m, n)
    U = DeviceTensor.empty(T.dims[:ncod] + (k,), ncod, T.ctx)
    S = DeviceTensor.empty((k,), 1, T.ctx)
    Vt = DeviceTensor.empty((k,) + T.dims[ncod:], 1, T.ctx)
    kk, eps = C.c_int64(), C.c_double()
    T.ctx.call("tnr_svd_trunc", T.ptr, len(T.dims), _lib.i64(T.dims), ncod, chi, U.ptr, S.ptr,
               Vt.ptr, C.byref(kk), C.byref(eps))
    assert kk.value == k
    return U, S, Vt, eps.value


def eigh_trunc(MM: DeviceTensor, chi: int):
    """eigh_trunc!(project_hermitian!(MM); trunc=truncrank(chi)) for a square matrix tensor."""
    nleg = len(MM.dims) // 2
    n = math.prod(MM.dims[:nleg])
    k = min(chi, n)
    W = DeviceTensor.empty((k,), 1, MM.ctx)
    V = DeviceTensor.empty(MM.dims[:nleg] + (k,), nleg, MM.ctx)
    kk, eps = C.c_int64(), C.c_double()
    MM.ctx.call("tnr_eigh_trunc", MM.ptr, n, chi, W.ptr, V.ptr, C.byref(kk), C.byref(eps))
    return W, V, eps.value
