"""Bit-level CPU model of the opt-in INT8 emulation engine (csrc/gemm_ozaki.cu) -- TEST
INFRASTRUCTURE ONLY.  Restates the arithmetic of `ozaki_split_kernel` / `ozaki_tile_kernel`:
power-of-two row scaling, signed base-128 digit planes with |q| <= 64, exact integer products of
the plane pairs with i + j = t, FP64 accumulation of the S groups from the smallest weight up.
"""
from __future__ import annotations

import numpy as np


def split(X, slices):
    """X: rows x K.  Returns digits[slices][rows][K] (int8 range) and scale[rows] = 2^e."""
    amax = np.abs(X).max(axis=1)
    # ilogb(amax) + 1: amax * 2^-e in [0.5, 1)  (frexp returns exactly that exponent)
    e = np.where(amax > 0, np.frexp(amax)[1], 0)
    rem = np.ldexp(X, (6 - e)[:, None]) / 128.0
    planes = []
    for _ in range(slices):
        y = rem * 128.0          # exact
        q = np.rint(y)           # |q| <= 64
        rem = y - q              # |rem| <= 0.5, exact
        planes.append(q.astype(np.int64))
    return planes, np.ldexp(1.0, e)


def reconstruct(planes, scale):
    """sum_i q_i 2^(-6-7i) * 2^e  (exact while the partial sums fit 53 bits)."""
    out = np.zeros(planes[0].shape)
    for i in reversed(range(len(planes))):
        out = out + np.ldexp(planes[i].astype(np.float64), -6 - 7 * i)
    return out * scale[:, None]


def multiply(A, B, slices=8):
    """C = A B^T with A: m x K, B: n x K, as the engine computes it."""
    pa, sa = split(A, slices)
    pb, sb = split(B, slices)
    C = np.zeros((A.shape[0], B.shape[0]))
    K = A.shape[1]
    assert slices * K * 4096 < 2 ** 31, "int32 accumulator would overflow"
    for t in reversed(range(slices)):            # smallest weight first
        acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for i in range(t + 1):
            acc += pa[i] @ pb[t - i].T           # exact integer product (int32 range on the GPU)
        assert np.abs(acc).max() < 2 ** 31
        C = C + acc.astype(np.float64) * (np.ldexp(1.0, -12 - 7 * t) * sa[:, None] * sb[None, :])
    return C


# ---- CRT variant ("ozaki_crt", csrc/crt_math.cuh + ozaki_crt_split_kernel / ozaki_tile_kernel<true>
# / ozaki_crt_reconstruct_kernel): bit-level model with exact Python integers for the CRT -----------
CRT_MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181]


def crt_bits(nmod, klog=14):
    import math

    P = math.prod(CRT_MODULI[:nmod])
    return min(53, (P.bit_length() - 2 - klog) // 2)


def crt_split(X, nmod):
    """X: rows x K.  residues[nmod][rows][K] in [-p/2, p/2) and scale[rows] = 2^(e - bits)."""
    bits = crt_bits(nmod)
    amax = np.abs(X).max(axis=1)
    e = np.where(amax > 0, np.frexp(amax)[1], 0)
    Xi = np.rint(np.ldexp(X, (bits - e)[:, None]))
    Xo = np.array([[int(v) for v in row] for row in Xi], dtype=object)
    res = []
    for p in CRT_MODULI[:nmod]:
        h = p // 2 if p % 2 == 0 else (p - 1) // 2
        res.append(np.array([[((v + h) % p) - h for v in row] for row in Xo], dtype=np.int64))
    return res, np.ldexp(1.0, e - bits), Xo


def multiply_crt(A, B, nmod=16):
    """C = A B^T as the CRT engine computes it (the reconstruction with exact integers; the FP64
    limb form of csrc/crt_math.cuh is checked against the same integers in
    tests/test_crt_math_host.py)."""
    import math

    ra, sa, _ = crt_split(A, nmod)
    rb, sb, _ = crt_split(B, nmod)
    ms = CRT_MODULI[:nmod]
    P = math.prod(ms)
    tot = np.zeros((A.shape[0], B.shape[0]), dtype=object)
    for i, p in enumerate(ms):
        acc = ra[i] @ rb[i].T                    # int32 range on the GPU
        assert np.abs(acc).max() < 2 ** 31
        w = (P // p) * pow(P // p, -1, p)
        tot = tot + (acc % p).astype(object) * w
    lift = np.vectorize(lambda x: ((x + P // 2) % P) - P // 2)(tot)
    return np.array(lift, dtype=np.float64) * sa[:, None] * sb[None, :]
