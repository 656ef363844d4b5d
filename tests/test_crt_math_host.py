"""Host check (g++, no GPU) of the scalar arithmetic of the CRT variant of the INT8 emulation
engine -- the very functions the kernels compile (`tnrkit.jl_b200/csrc/crt_math.cuh`: residues of
the scaled operands, residues of the INT32 accumulators, FP64-limb reconstruction) -- against
exact Python integers, and of the generated constant tables against tools/gen_crt_tables.py."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("crt") / "libcrt_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "crt_host_shim.cpp")], check=True)
    return C.CDLL(so)


def test_crt_tables_are_current():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import contextlib
    import io

    import gen_crt_tables

    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gen_crt_tables.main()
    assert buf.getvalue() == open(os.path.join(ROOT, "tnrkit.jl_b200", "csrc", "crt_tables.inc")).read()
    for nmod in range(14, 19):
        t = gen_crt_tables.table(nmod)
        assert math.prod(t["p"]) > 2 * 2 ** 14 * 2 ** (2 * t["bits"])
    assert gen_crt_tables.table(16)["bits"] == 53


@pytest.mark.parametrize("nmod", [14, 16, 18])
@pytest.mark.parametrize("K", [512, 13824])
def test_crt_host_arithmetic_is_exact(shim, nmod, K):
    rng = np.random.default_rng(nmod * 7 + K)
    m, n = 6, 5
    A = rng.standard_normal((m, K)) * np.exp(rng.uniform(-8, 8, size=(m, K)))
    B = rng.standard_normal((n, K)) * np.exp(rng.uniform(-8, 8, size=(n, 1)))
    A[0, :7] = 0.0
    B[1] = 0.0                                   # an all-zero row: exponent 0, residues 0

    def split(X):
        rows = X.shape[0]
        out = np.zeros((nmod, rows, K), dtype=np.int8)
        scale = np.zeros(rows)
        bits = C.c_int()
        rc = shim.crt_host_split(np.ascontiguousarray(X).ctypes.data_as(C.c_void_p),
                                 C.c_longlong(rows), C.c_longlong(K), nmod,
                                 out.ctypes.data_as(C.c_void_p), scale.ctypes.data_as(C.c_void_p),
                                 C.byref(bits))
        assert rc == 0
        return out, scale, bits.value

    ra, sa, bits = split(A)
    rb, sb, _ = split(B)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_crt_tables

    ps = gen_crt_tables.table(nmod)["p"]
    # residues equal the exact symmetric residues of the scaled integers
    Ai = [[int(np.rint(np.ldexp(x, -int(np.log2(sa[r])))) ) for x in A[r]] for r in range(m)]
    Bi = [[int(np.rint(np.ldexp(x, -int(np.log2(sb[r])))) ) for x in B[r]] for r in range(n)]
    assert max(abs(v) for row in Ai for v in row) <= 2 ** bits
    for i, p in enumerate(ps):
        for r in (0, m - 1):
            want = np.array([((v + p // 2) % p) - p // 2 if p % 2 == 0 else ((v + (p - 1) // 2) % p) - (p - 1) // 2
                             for v in Ai[r]])
            assert np.array_equal(ra[i, r].astype(np.int64), want), (p, r)
    assert np.abs(ra.astype(np.int64)).max() <= 128 and ra.min() >= -128 and ra.max() <= 127
    # INT32 accumulators of the residue GEMMs (what tcgen05 kind::i8 produces)
    acc = np.stack([ra[i].astype(np.int64) @ rb[i].astype(np.int64).T for i in range(nmod)])
    assert np.abs(acc).max() < 2 ** 31
    acc32 = np.ascontiguousarray(acc.astype(np.int32)).reshape(nmod, m * n)
    out = np.zeros(m * n)
    assert shim.crt_host_reconstruct(acc32.ctypes.data_as(C.c_void_p), C.c_longlong(m * n), nmod,
                                     out.ctypes.data_as(C.c_void_p)) == 0
    exact = [[sum(x * y for x, y in zip(Ai[i], Bi[j])) for j in range(n)] for i in range(m)]
    for i in range(m):
        for j in range(n):
            got, want = out[i * n + j], exact[i][j]
            assert abs(int(got) - want) <= max(1, abs(want)) * 2.3e-16, (i, j, got, want)
    # end to end: FP64-level accuracy of the emulated product
    Cemu = out.reshape(m, n) * sa[:, None] * sb[None, :]
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    mag = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    err = float(np.max(np.abs(Cemu - ref) / np.where(mag > 0, mag, 1)))
    assert err <= {14: 2e-13, 16: 3e-16, 18: 3e-16}[nmod], err


def test_bit_level_model_agrees_with_host_arithmetic(shim):
    """oracle/ozaki_model.py: multiply_crt (exact-integer CRT) == the FP64-limb pipeline of
    crt_math.cuh on the same operands, to the final rounding."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ozaki_model

    rng = np.random.default_rng(5)
    m, n, K, nmod = 4, 3, 2048, 16
    A = rng.standard_normal((m, K)) * np.exp(rng.uniform(-6, 6, size=(m, K)))
    B = rng.standard_normal((n, K))
    want = ozaki_model.multiply_crt(A, B, nmod)
    ra, sa, _ = ozaki_model.crt_split(A, nmod)
    rb, sb, _ = ozaki_model.crt_split(B, nmod)
    acc = np.ascontiguousarray(np.stack([ra[i] @ rb[i].T for i in range(nmod)]).astype(np.int32)
                               ).reshape(nmod, m * n)
    out = np.zeros(m * n)
    assert shim.crt_host_reconstruct(acc.ctypes.data_as(C.c_void_p), C.c_longlong(m * n), nmod,
                                     out.ctypes.data_as(C.c_void_p)) == 0
    got = out.reshape(m, n) * sa[:, None] * sb[None, :]
    assert np.max(np.abs(got - want) / np.abs(want)) <= 2.3e-16
