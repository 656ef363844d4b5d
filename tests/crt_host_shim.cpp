// Host build (g++) of csrc/crt_math.cuh for tests/test_crt_math_host.py -- test infrastructure.
#include "../tnrkit.jl_b200/csrc/crt_math.cuh"

extern "C" {
// rows x K operand (row major here: row r contiguous) -> residues[nmod][rows][K], scale[rows]
int crt_host_split(const double* X, long long rows, long long K, int nmod, signed char* out,
                   double* scale, int* bits_out) {
    if (nmod < CRT_MIN_MOD || nmod > CRT_MAX_MOD) return 1;
    const CrtTable& t = CRT_TABLES[nmod - CRT_MIN_MOD];
    *bits_out = t.bits;
    for (long long r = 0; r < rows; ++r) {
        double amax = 0.0;
        for (long long k = 0; k < K; ++k) amax = std::fmax(amax, std::fabs(X[r * K + k]));
        int e = 0;
        if (amax > 0.0 && std::isfinite(amax)) e = std::ilogb(amax) + 1;
        scale[r] = std::ldexp(1.0, e - t.bits);
        for (long long k = 0; k < K; ++k) {
            const double Xi = std::rint(std::ldexp(X[r * K + k], t.bits - e));
            for (int i = 0; i < nmod; ++i)
                out[((long long)i * rows + r) * K + k] =
                    (signed char)tnr::crt_residue(Xi, t.p[i], t.inv_p[i]);
        }
    }
    return 0;
}
// acc[nmod][n] int32 accumulators -> out[n] reconstructed integers (as doubles)
int crt_host_reconstruct(const int* acc, long long n, int nmod, double* out) {
    if (nmod < CRT_MIN_MOD || nmod > CRT_MAX_MOD) return 1;
    const CrtTable& t = CRT_TABLES[nmod - CRT_MIN_MOD];
    unsigned char res[CRT_MAX_MOD];
    for (long long e = 0; e < n; ++e) {
        for (int i = 0; i < nmod; ++i)
            res[i] = (unsigned char)tnr::crt_acc_residue(acc[(long long)i * n + e], t.p[i], t.inv_p[i]);
        out[e] = tnr::crt_reconstruct(res, 1, t);
    }
    return 0;
}
}
