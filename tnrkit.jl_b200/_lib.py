"""ctypes binding of libtnrcuda.so (include/tnrcuda.h).

This is the Python twin of the `ccall` stubs a Julia maintainer would add (see
INTEGRATION.md).  There is no CPU fallback: if the shared library is missing, or no
CUDA device is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtnrcuda.so")

TNR_TRG, TNR_BTRG, TNR_HOTRG, TNR_ATRG, TNR_HOTRG_3D, TNR_ATRG_3D = range(6)

_c_i64p = C.POINTER(C.c_int64)
_c_dp = C.c_void_p  # raw device pointers travel as integers

# name -> (argtypes)  ; every function returns int except where noted
class GemmProblem(C.Structure):
    """tnr_gemm_problem"""
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("A", C.c_void_p),
                ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64), ("C", C.c_void_p),
                ("ldc", C.c_int64)]


_SIGNATURES = {
    "tnr_version": [],
    "tnr_create": [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)],
    "tnr_destroy": [C.c_void_p],
    "tnr_last_error": [C.c_void_p],
    "tnr_synchronize": [C.c_void_p],
    "tnr_get_counters": [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                         C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "tnr_reset_counters": [C.c_void_p],
    "tnr_get_tma_launches": [C.c_void_p, C.POINTER(C.c_uint64)],
    "tnr_set_option": [C.c_void_p, C.c_char_p, C.c_int64],
    "tnr_get_counter": [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)],
    "tnr_gemm_timing": [C.c_void_p, C.c_int],
    "tnr_gemm_timing_read": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), _c_i64p],
    "tnr_malloc": [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)],
    "tnr_free": [C.c_void_p, C.c_void_p],
    "tnr_upload": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
    "tnr_download": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
    "tnr_gemm": [C.c_void_p, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, _c_dp,
                 C.c_int64, _c_dp, C.c_int64, C.c_double, _c_dp, C.c_int64],
    "tnr_gemm_strided_batched": [C.c_void_p, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int,
                                 C.c_double, _c_dp, C.c_int64, C.c_int64, _c_dp, C.c_int64,
                                 C.c_int64, C.c_double, _c_dp, C.c_int64, C.c_int64, C.c_int],
    "tnr_gemm_ozaki": [C.c_void_p, C.c_int, C.c_int, C.c_int, _c_dp, C.c_int64, _c_dp, C.c_int64,
                       _c_dp, C.c_int64],
    "tnr_gemm_grouped": [C.c_void_p, C.c_char, C.c_char, C.c_int, C.POINTER(GemmProblem),
                         C.c_double, C.c_double],
    "tnr_permute": [C.c_void_p, _c_dp, _c_dp, C.c_int, _c_i64p, C.POINTER(C.c_int)],
    "tnr_strided_copy": [C.c_void_p, _c_dp, _c_dp, C.c_int, _c_i64p, _c_i64p, _c_i64p],
    "tnr_strided_sum": [C.c_void_p, _c_dp, C.c_int, _c_i64p, _c_i64p, C.POINTER(C.c_void_p),
                        C.POINTER(C.c_double)],
    "tnr_scale": [C.c_void_p, _c_dp, C.c_int64, C.c_double],
    "tnr_diag_scale": [C.c_void_p, _c_dp, C.c_int64, C.c_int64, C.c_int64, _c_dp, C.c_int, C.c_int,
                       C.c_double],
    "tnr_axis_scale": [C.c_void_p, _c_dp, C.c_int64, C.c_int64, C.c_int64, _c_dp, C.c_int,
                       C.c_double],
    "tnr_vec_map": [C.c_void_p, _c_dp, _c_dp, C.c_int64, C.c_int, C.c_double],
    "tnr_topk_select": [C.c_void_p, _c_dp, C.c_int64, C.c_int64, C.POINTER(C.c_int32),
                        C.POINTER(C.c_double)],
    "tnr_contract": [C.c_void_p, _c_dp, C.c_int, _c_i64p, C.c_char_p, _c_dp, C.c_int, _c_i64p,
                     C.c_char_p, _c_dp, C.c_char_p],
    "tnr_svd_trunc": [C.c_void_p, _c_dp, C.c_int, _c_i64p, C.c_int, C.c_int, _c_dp, _c_dp, _c_dp,
                      _c_i64p, C.POINTER(C.c_double)],
    "tnr_orth_r": [C.c_void_p, _c_dp, C.c_int, _c_i64p, C.c_int, _c_dp],
    "tnr_psd_factor": [C.c_void_p, _c_dp, C.c_int64, _c_dp, _c_i64p],
    "tnr_fill_random": [C.c_void_p, _c_dp, C.c_int64, C.c_uint64],
    "tnr_orthonormalize": [C.c_void_p, _c_dp, C.c_int64, C.c_int64, C.POINTER(C.c_int)],
    "tnr_qr": [C.c_void_p, _c_dp, C.c_int64, C.c_int64, _c_dp, _c_dp],
    "tnr_eigh_trunc": [C.c_void_p, _c_dp, C.c_int64, C.c_int, _c_dp, _c_dp, _c_i64p,
                       C.POINTER(C.c_double)],
    "tnr_step_out_dims": [C.c_int, _c_i64p, C.c_int, _c_i64p],
    "tnr_trg_step": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p],
    "tnr_btrg_step": [C.c_void_p, _c_dp, _c_i64p, _c_dp, _c_dp, C.c_double, C.c_int, _c_dp,
                      _c_i64p, _c_dp, _c_dp],
    "tnr_hotrg_step": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p],
    "tnr_atrg_step": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p],
    "tnr_hotrg3d_step": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p],
    "tnr_hotrg3d_substep": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p, C.c_int64,
                            C.c_int64],
    "tnr_hotrg3d_substep_peers": [C.c_void_p, _c_dp, _c_i64p, C.c_int, C.POINTER(C.c_void_p),
                                  C.c_int, C.c_int, _c_i64p, C.c_int64, C.c_int64],
    "tnr_hotrg3d_proj_half": [C.c_void_p, _c_dp, _c_i64p, C.c_int, C.c_int, _c_dp],
    "tnr_hotrg3d_contract": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, C.POINTER(C.c_void_p),
                             C.c_int, C.c_int, _c_i64p, C.c_int64, C.c_int64],
    "tnr_atrg3d_step": [C.c_void_p, _c_dp, _c_i64p, C.c_int, _c_dp, _c_i64p],
    "tnr_finalize_2d": [C.c_void_p, _c_dp, _c_i64p, C.POINTER(C.c_double)],
    "tnr_finalize_btrg": [C.c_void_p, _c_dp, _c_i64p, _c_dp, _c_dp, C.POINTER(C.c_double)],
    "tnr_finalize_3d": [C.c_void_p, _c_dp, _c_i64p, C.POINTER(C.c_double)],
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class TNRCudaError(RuntimeError):
    pass


def load():
    """Loads libtnrcuda.so (built by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TNRCudaError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (tnrkit.jl_b200 has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "tnr_last_error" else C.c_int
    _lib = lib
    return lib


def i64(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr


def i32(values):
    return (C.c_int * len(values))(*[int(v) for v in values])


class Context:
    """One engine context per process/GPU (one process per GPU under torchrun)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.tnr_create(int(device), C.c_void_p(stream or 0), C.byref(h))
        if rc != 0:
            msg = self.lib.tnr_last_error(None)
            raise TNRCudaError(f"tnr_create failed ({rc}): {msg.decode() if msg else ''}")
        self.h = h
        self.device = device

    @property
    def torch_device(self) -> str:
        """Device string of the buffers this context's kernels may touch (always CUDA)."""
        return f"cuda:{self.device}"

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.tnr_last_error(self.h)
            raise TNRCudaError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def call(self, name: str, *args):
        self.check(getattr(self.lib, name)(self.h, *args), name)

    def synchronize(self):
        self.call("tnr_synchronize")

    def counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        f, pb = C.c_double(), C.c_double()
        self.call("tnr_get_counters", C.byref(a), C.byref(b), C.byref(f), C.byref(pb))
        t = C.c_uint64()
        self.call("tnr_get_tma_launches", C.byref(t))
        g = C.c_double()
        self.call("tnr_get_counter", b"grouped_gemm_launches", C.byref(g))
        return {"launches": a.value, "gemm_launches": b.value, "gemm_flops": f.value,
                "permute_bytes": pb.value, "tma_gemm_launches": t.value,
                "grouped_gemm_launches": int(g.value)}

    def set_option(self, key: str, value: int):
        self.call("tnr_set_option", key.encode(), int(value))

    def reset_counters(self):
        self.call("tnr_reset_counters")

    def gemm_timing(self, enable: bool):
        self.call("tnr_gemm_timing", 1 if enable else 0)

    def gemm_timing_read(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self.call("tnr_gemm_timing_read", C.byref(ms), C.byref(fl), C.byref(n))
        return ms.value, fl.value, n.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.tnr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context() -> Context:
    """Context on the current torch CUDA device and its current stream."""
    global _default_ctx
    if _default_ctx is None:
        import torch

        if not torch.cuda.is_available():
            raise TNRCudaError("no CUDA device: tnrkit.jl_b200 has no CPU fallback")
        dev = torch.cuda.current_device()
        stream = torch.cuda.current_stream().cuda_stream
        _default_ctx = Context(dev, stream)
    return _default_ctx
