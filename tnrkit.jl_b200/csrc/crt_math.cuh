// Scalar arithmetic of the CRT variant of the INT8 emulation engine (gemm_ozaki.cu, option
// "ozaki_crt"): residues of a scaled FP64 operand, residue of an INT32 accumulator, and the
// FP64-only reconstruction of the exact integer dot product from its residues.  Plain
// host/device inline functions so that the same code is compiled into the kernels and into the
// host-side check of tests/test_crt_math_host.py (g++, no GPU).
//
// Scheme (Ozaki scheme II): X = rint(x 2^(bits - e_row)) is an integer of at most `bits` bits;
// for pairwise coprime p_1..p_N <= 256 with P = prod p_i > 2 K 2^(2 bits) the integer dot product
// sum_k X_k Y_k is determined by its residues mod p_i, and each residue is ONE int8 x int8 ->
// int32 dot product of the symmetric residues (|X mod p| <= 128, K 128^2 < 2^31).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define TNR_HD __host__ __device__ __forceinline__
#else
#define TNR_HD inline
#endif

#include "crt_tables.inc"

namespace tnr {

// symmetric residue of the integer-valued double X (|X| <= 2^53) in [-p/2, p/2): fits int8
TNR_HD int crt_residue(double X, int p, double inv_p) {
    const double q = rint(X * inv_p);            // off by at most one from the exact quotient
    int r = (int)fma(-q, (double)p, X);          // exact: the true value is an integer |r| < 2p
    if (2 * r >= p) r -= p;
    else if (2 * r < -p) r += p;
    return r;
}

// residue in [0, p) of an int32 accumulator (|v| < 2^31)
TNR_HD int crt_acc_residue(int v, int p, double inv_p) {
    const int q = (int)rint((double)v * inv_p);
    int r = v - q * p;                           // |r| <= p/2 + 1
    if (r < 0) r += p;
    if (r >= p) r -= p;
    return r;
}

// quotient q = rint(x / P) from the leading limb sums, then Horner over the corrected limbs;
// NLV (number of limbs incl. one spare) is a template parameter so that S stays in registers
template <int NLV>
TNR_HD double crt_finish(const double* S, const CrtTable& t) {
    const double top = S[NLV - 1] * 4294967296.0 + S[NLV - 2] + S[NLV - 3] * (1.0 / 4294967296.0);
    const double q = rint(top / t.Pscaled);      // |q| <= 256 N
    double out = 0.0;
#pragma unroll
    for (int j = NLV - 1; j >= 0; --j) out = out * 4294967296.0 + (S[j] - q * t.PL[j]);
    return out;
}

// symmetric lift of sum_i res[i] w_i mod P as a double; res[i] in [0, p_i), stride between
// consecutive moduli = `stride` bytes.  Every product and sum below is exact in FP64 (limb sums
// < 2^45); only the final Horner sum rounds (relative 2^-53 of the value itself).
TNR_HD double crt_reconstruct(const unsigned char* res, long long stride, const CrtTable& t) {
    double S[CRT_NL];
#pragma unroll
    for (int j = 0; j < CRT_NL; ++j) S[j] = 0.0;
    for (int i = 0; i < t.nmod; ++i) {
        const double r = (double)res[(long long)i * stride];
#pragma unroll
        for (int j = 0; j < CRT_NL; ++j) S[j] += r * t.W[i][j];
    }
    return t.nl == 5 ? crt_finish<5>(S, t) : crt_finish<6>(S, t);
}

}  // namespace tnr
