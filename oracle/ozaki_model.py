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
