"""CPU prototype (numpy + Python integers) of the CRT variant of the Ozaki scheme -- groundwork
for round 2, NOT part of the product.

Idea (Ozaki/Uchino/Imamura, "Ozaki scheme II"): scale the rows of A and the columns of B to
integers of `bits` bits, A' = rint(2^bits * A / rowmax), B' likewise.  The exact integer product
C' = A' B'^T is bounded by K * 2^(2*bits).  Pick pairwise coprime moduli p_1..p_N <= 256 with
P = prod p_i > 2 max|C'|; then C' is determined by its residues C_i = (A' mod p_i)(B' mod p_i)^T
mod p_i, and each residue product is ONE int8 x int8 -> int32 GEMM (|A' mod p| <= 128,
K * 128^2 < 2^31).  N ~ (2*bits + log2 K + 2) / 8 GEMMs instead of S(S+1)/2 = 36 for the
digit-plane scheme with S = 8.

This script measures, for the chunk-GEMM shape statistics (K = 13824), how many moduli are
needed for a given `bits` and what accuracy results, using exact Python integers for the CRT.
"""
import math
import sys

import numpy as np

MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191,
          181, 179, 173, 167, 163, 157, 151, 149]


def pairwise_coprime(ms):
    return all(math.gcd(a, b) == 1 for i, a in enumerate(ms) for b in ms[i + 1:])


def emulate(A, B, bits, nmod):
    """A: m x K, B: n x K (rows are the K-major operands).  Returns C ~ A B^T."""
    ms = MODULI[:nmod]
    assert pairwise_coprime(ms)
    P = math.prod(ms)
    ea = np.ceil(np.log2(np.abs(A).max(axis=1)))
    eb = np.ceil(np.log2(np.abs(B).max(axis=1)))
    Ai = np.rint(np.ldexp(A, (bits - ea).astype(int)[:, None])).astype(object)  # |Ai| <= 2^bits
    Bi = np.rint(np.ldexp(B, (bits - eb).astype(int)[:, None])).astype(object)
    K = A.shape[1]
    bound = K * (1 << (2 * bits))
    assert P > 2 * bound, (math.log2(P), math.log2(2 * bound))
    # residue GEMMs (these are the int8 tensor-core products on the GPU)
    C = np.zeros((A.shape[0], B.shape[0]), dtype=object)
    for p in ms:
        ar = np.array([[int(x) % p for x in row] for row in Ai], dtype=np.int64)
        br = np.array([[int(x) % p for x in row] for row in Bi], dtype=np.int64)
        ar = np.where(ar > p // 2, ar - p, ar)   # symmetric residues fit int8
        br = np.where(br > p // 2, br - p, br)
        assert np.abs(ar).max() <= 128 and np.abs(br).max() <= 128
        ci = (ar @ br.T) % p                      # int32-exact on the GPU: K * 128^2 < 2^31
        w = (P // p) * pow(P // p, -1, p)         # CRT weight
        C = C + ci.astype(object) * w
    C = np.vectorize(lambda x: ((x + P // 2) % P) - P // 2)(C)   # symmetric lift
    scale = np.ldexp(1.0, (ea[:, None] + eb[None, :] - 2 * bits).astype(int))
    return np.array(C, dtype=np.float64) * scale, len(ms), math.log2(P)


def main():
    rng = np.random.default_rng(0)
    m = n = 24
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 13824
    A = rng.standard_normal((m, K)) * np.exp(rng.uniform(-6, 6, size=(m, K)))
    B = rng.standard_normal((n, K)) * np.exp(rng.uniform(-6, 6, size=(n, 1)))
    ref = A.astype(np.longdouble) @ B.T.astype(np.longdouble)
    scale = np.abs(A).astype(np.longdouble) @ np.abs(B.T).astype(np.longdouble)
    dgemm = A @ B.T
    print(f"K = {K};  DGEMM max |err|/(|A||B|) = {float(np.max(np.abs(dgemm - ref) / scale)):.2e}")
    for bits in (40, 46, 50, 53, 56):
        need = 2 * bits + math.ceil(math.log2(K)) + 2
        nmod = next(k for k in range(1, len(MODULI) + 1) if math.log2(math.prod(MODULI[:k])) > need)
        C, k, lp = emulate(A, B, bits, nmod)
        err = float(np.max(np.abs(C - ref) / scale))
        print(f"bits = {bits}: {k} moduli (log2 P = {lp:.1f} > {need}) -> {k} int8 GEMMs, "
              f"max |err|/(|A||B|) = {err:.2e}")


if __name__ == "__main__":
    main()


# ----------------------------------------------------------------------------------------------
# FP64-only CRT reconstruction (what the GPU epilogue / reconstruction kernel would execute):
# no big integers, only exactly-representable double arithmetic on 32-bit limbs.
# ----------------------------------------------------------------------------------------------
def limbs32(x, n):
    return [float((x >> (32 * j)) & 0xFFFFFFFF) for j in range(n)]


def reconstruct_fp64(residues, ms):
    """residues[i]: float64 array of c_i in [0, p_i).  Returns the symmetric lift of
    sum_i c_i w_i mod P as float64 with absolute accuracy ~ P * 2^-60."""
    P = math.prod(ms)
    nl = (P.bit_length() + 31) // 32 + 1
    W = [limbs32((P // p) * pow(P // p, -1, p), nl) for p in ms]
    PL = limbs32(P, nl)
    S = [sum(residues[i] * W[i][j] for i in range(len(ms))) for j in range(nl)]  # exact: < 2^44
    # quotient q = round(x / P) from the two most significant non-zero limbs (|q| <= 256 N)
    top = S[nl - 1] * 2.0 ** 32 + S[nl - 2] + S[nl - 3] * 2.0 ** -32
    q = np.rint(top / (float(P) / 2.0 ** (32 * (nl - 2))))
    R = [S[j] - q * PL[j] for j in range(nl)]                                    # exact: < 2^45
    out = np.zeros_like(residues[0])
    for j in reversed(range(nl)):
        out = out * 2.0 ** 32 + R[j]       # Horner in FP64: error ~ 2^-53 of the partial sums
    return out


def check_reconstruction():
    rng = np.random.default_rng(1)
    ms = MODULI[:16]
    P = math.prod(ms)
    n = 4000
    # random integers |x| < P/4 with a wide range of magnitudes
    xs = [int(rng.integers(-2 ** 62, 2 ** 62)) * (1 << int(rng.integers(0, 60))) for _ in range(n)]
    xs = [x % P if abs(x) < P // 4 else (x % (P // 4)) for x in xs]
    xs = [x - P if x > P // 2 else x for x in xs]
    res = [np.array([x % p for x in xs], dtype=np.float64) for p in ms]
    got = reconstruct_fp64(res, ms)
    err = max(abs(float(g) - float(x)) for g, x in zip(got, xs))
    rel = max(abs(int(g) - x) / max(1, abs(x)) for g, x in zip(got, xs))
    print(f"FP64 limb reconstruction: max abs error / P = {err / float(P):.2e}; max error relative "
          f"to the value itself = {rel:.2e} (FP64 rounding of the exact integer is 1.1e-16)")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "recon":
    check_reconstruction()
