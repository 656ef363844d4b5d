// Common definitions for libtnrcuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace tnr {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define TNR_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            throw ::tnr::Error(2, std::string(#expr) + ": " + cudaGetErrorString(_e) +   \
                                      " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

#define TNR_CHECK(cond, msg)                                                             \
    do {                                                                                 \
        if (!(cond))                                                                     \
            throw ::tnr::Error(1, std::string(msg) + " [" #cond "] at " + __FILE__ +     \
                                      ":" + std::to_string(__LINE__));                   \
    } while (0)

// Counters the bench reads back: how many of OUR kernels were launched.
struct Counters {
    unsigned long long launches = 0;       // all kernels of this library
    unsigned long long gemm_launches = 0;  // DMMA GEMM kernels
    unsigned long long grouped_gemm_launches = 0;  // of which grouped (per-sector) launches
    unsigned long long tma_gemm_launches = 0;  // of which TMA/mbarrier warp-specialised
    unsigned long long ozaki_launches = 0;   // INT8 tcgen05 group kernels
    unsigned long long ozaki_gemms = 0;      // FP64-equivalent GEMMs served by the Ozaki engine
    unsigned long long tma_grouped_launches = 0;   // grouped (per-sector) launches of the TMA GEMM
    unsigned long long peer_scatter_launches = 0;  // slab scatters that also stored to peer GPUs
    unsigned long long permute_bulk_launches = 0;  // strided copies served by the TMA-fed kernel
    unsigned long long preconditioned_jacobi = 0;  // tall Jacobi problems preconditioned by Gram eigenvectors
    unsigned long long subspace_eigh = 0;       // eigh_trunc calls served by the subspace solver
    unsigned long long subspace_svd = 0;        // svd_trunc calls served by the subspace solver
    unsigned long long persistent_jacobi = 0;    // Jacobi iterations run as one cooperative launch
    unsigned long long qr_factorizations = 0;    // blocked Householder QR factorizations (qr.cu)
    unsigned long long psd_factorizations = 0;   // pivoted Cholesky factorizations (pchol.cu)
    unsigned long long cholqr2 = 0;              // CholeskyQR2 orthonormalisations (pchol.cu)
    unsigned long long cholqr2_refused = 0;      // ... refused (ill-conditioned): Householder path taken
    unsigned long long jacobi_limit_accepted = 0; // sweep limit reached with only rounding-level rotations left
    unsigned long long jacobi_not_converged = 0; // one-sided Jacobi runs that hit the sweep limit (an error is raised)
    unsigned long long subspace_fallbacks = 0;  // ... that fell back to full Jacobi
    double gemm_flops = 0.0;               // 2*m*n*k summed over GEMM launches
    double permute_bytes = 0.0;            // read+write bytes moved by permute kernels
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int num_sms = 148;
    Counters ctr;
    // optional timing of the big GEMM (dominant kernel) with events on ctx->stream
    bool time_gemm = false;
    bool disable_tma = false;  // force the cp.async GEMM (A/B testing)
    bool disable_subspace = false;  // force full Jacobi in eigh_trunc
    bool disable_block_jacobi = false;  // force the one-pair-per-CTA Jacobi rounds
    int permute_tile = 96;   // composite run length of the tiled permute kernel (opt-in: 32 | 48 | 64)
    int permute_unroll = 4;  // 1 | 2 | 4: rows of the read phase in flight per thread in the fallback tiled kernel (r02: 4 measured 1.7x faster than 1)
    int permute_tpc = 8;       // max tiles per CTA of the TMA-fed copy (tuning)
    bool permute_dense = true;  // dense fetch + in-place re-pitch of short contiguous source rows
    int permute_chunk_below = 0;  // rows below this many bytes: cp.async chunks instead of bulk pieces (tuning)
    bool permute_bulk = true;  // TMA-fed tiled copy (cp.async.bulk) whenever the source pieces are 16-byte aligned
    int ozaki_crt = 0;     // 14..18: CRT variant of the INT8 engine with that many moduli (opt-in, unmeasured)
    int ozaki_slices = 0;  // > 0: INT8 Ozaki engine for the big TN contractions (opt-in)
    bool disable_precondition = false;  // no Gram preconditioning of tall Jacobi problems
    int jacobi_max_sweeps = 40;  // a Jacobi iteration that needs more raises an error
    bool disable_persistent_jacobi = false;  // one launch per Jacobi round + host sync per sweep (round 1)
    bool disable_cholqr = false;  // subspace bases: Householder QR instead of CholeskyQR2 (A/B tests)
    int jacobi_max_bc = 8;  // largest column block of the shared-memory Jacobi (16 | 8 | 4); 8: twice the CTAs of 16 and one pass of the 8 warps per inner round
    bool disable_qr = false;  // tall problems: Gram-preconditioned Jacobi (round 1) instead of Householder QR + Jacobi of R
    double timed_flops = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
    // optional per-phase timing of the scheme bodies (same switch as time_gemm): name + events
    struct PhaseEvt { const char* name; cudaEvent_t e0, e1; };
    std::vector<PhaseEvt> phase_events;
    long long hotrg3d_pk_budget = 48LL << 30;  // bytes of absorbed operands Pk_d held at once
    std::string last_error;
    // multi-GPU sharding of the HOTRG_3D open bond (set by tnr_set_shard)
    int rank = 0, world = 1;
};

// Times a phase of a scheme body with CUDA events on ctx->stream while gemm timing is on
// (bench.py); free otherwise.  Read back with tnr_get_counter("phase_ms.<name>").
struct PhaseScope {
    Context* ctx;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const char* name;
    PhaseScope(Context* c, const char* nm) : ctx(c), name(nm) {
        if (!ctx->time_gemm) return;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
    }
    void stop() {
        if (!e0 || !e1) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->phase_events.push_back({name, e0, e1});
        e0 = e1 = nullptr;
    }
    ~PhaseScope() { stop(); }
};

// ---- device memory (stream ordered) ----
double* dalloc(Context* ctx, size_t n_doubles);
void dfree(Context* ctx, void* p);

// ---- kernels: gemm_dmma.cu ----
// C(m x n, ldc) = alpha * op(A) * op(B) + beta * C, column major, FP64 tensor cores.
// Batched with two batch levels: batch index b = b1 + nb1*b2 (b1 < nb1, b2 < nb2),
// operand offset = b1*s1 + b2*s2 (in doubles).
struct GemmBatch {
    int nb1 = 1, nb2 = 1;
    long long sA1 = 0, sA2 = 0, sB1 = 0, sB2 = 0, sC1 = 0, sC2 = 0;
};
void gemm(Context* ctx, char transa, char transb, int m, int n, int k, double alpha,
          const double* A, long long lda, const double* B, long long ldb, double beta,
          double* C, long long ldc, const GemmBatch& batch = GemmBatch());

// grouped launch: independent problems (per-sector blocks) sharing op(A), op(B), alpha, beta
struct GroupedProblem {
    int m, n, k;
    const double* A;
    long long lda;
    const double* B;
    long long ldb;
    double* C;
    long long ldc;
};
void gemm_grouped(Context* ctx, char transa, char transb, const std::vector<GroupedProblem>& probs,
                  double alpha, double beta);
bool gemm_grouped_tma_tn(Context* ctx, const std::vector<GroupedProblem>& probs, double alpha,
                         double beta);

// gemm_tma.cu: warp-specialised TMA + mbarrier + DMMA kernel for the TN layout; returns false
// when the problem does not fit (caller falls back to the cp.async kernel)
bool gemm_tma_tn(Context* ctx, int m, int n, int k, double alpha, const double* A, long long lda,
                 const double* B, long long ldb, double beta, double* C, long long ldc);
// grouped TN products (per-sector blocks) on the same kernel, one launch; false = not applicable

// gemm_ozaki.cu: FP64-accurate TN GEMM on the INT8 tensor cores (tcgen05 + TMEM), opt-in
struct OzakiOperand {        // int8 digit planes of the rows of a K-major FP64 matrix
    int8_t* planes = nullptr;  // [slices][rows][K]
    double* scale = nullptr;   // [rows] power-of-two row scale
    long long rows = 0, K = 0;
    int slices = 0;            // digit planes, or moduli when crt
    bool crt = false;          // planes are residues modulo the first `slices` CRT moduli
};
bool ozaki_applicable(const Context* ctx, long long m, long long n, long long k);
OzakiOperand ozaki_split(Context* ctx, const double* X, long long ld, long long rows, long long K);
void ozaki_free(Context* ctx, OzakiOperand& o);
void ozaki_multiply(Context* ctx, const OzakiOperand& A, const OzakiOperand& B, double* C,
                    long long ldc);

// ---- kernels: permute.cu ----
// dst[sum i_j*dstride_j] = src[sum i_j*sstride_j] for all multi-indices i < dims.
void strided_copy(Context* ctx, const double* src, double* dst, int rank,
                  const long long* dims, const long long* sstride, const long long* dstride);
void strided_copy_multi(Context* ctx, const double* src, double* const* dsts, int ndst, int rank,
                        const long long* dims, const long long* sstride, const long long* dstride);
// dst (compact, column major, dims[perm[k]]) leg k = src leg perm[k]
void permute(Context* ctx, const double* src, double* dst, int rank, const long long* dims,
             const int* perm);

// ---- kernels: elementwise.cu ----
void scale(Context* ctx, double* x, long long n, double alpha);
// x[i] *= 1 / *dev_scalar
void scale_inv_dev(Context* ctx, double* x, long long n, const double* dev_scalar);
// A(m x n, lda): A[:,j] *= f(s[j]) (cols) or A[i,:] *= f(s[i]) (rows)
// mode: 0 identity, 1 sqrt, 2 pseudopow(s, p)
void diag_scale(Context* ctx, double* A, long long m, long long n, long long lda,
                const double* s, bool rows, int mode, double p);
void vec_map(Context* ctx, const double* s, double* out, long long n, int mode, double p);
// generic strided trace-like reduction: out = | sum_i x[off + i*stride pattern] |
// sum over multi-index i<dims of src[sum i_j*stride_j]; result written to dev out (abs if absval)
void strided_sum(Context* ctx, const double* src, int rank, const long long* dims,
                 const long long* stride, double* dev_out, bool absval);
void strided_sum_w(Context* ctx, const double* src, int rank, const long long* dims,
                   const long long* stride, const double* const* weights, double* dev_out,
                   bool absval);
// A viewed as [m1][n][m2] (column major): A[i,j,k] *= f(s[j])
void axis_scale(Context* ctx, double* A, long long m1, long long n, long long m2, const double* s,
                int mode, double p);
void symmetrize(Context* ctx, double* A, long long n);  // A = (A + A^T)/2
void set_identity(Context* ctx, double* A, long long n);
void fill_zero(Context* ctx, double* A, long long n);
void fill_random(Context* ctx, double* x, long long n, unsigned long long seed);
void axpy(Context* ctx, double* y, const double* x, double alpha, long long n);  // y += alpha x
void sum_squares(Context* ctx, const double* x, long long n, double* dev_out);   // deterministic
void sqrt_inplace(Context* ctx, const double* dev_in, double* dev_out);
// A (m x n, lda): column j <- 0 where |vals[j]| <= rel * |vals[0]| (vals sorted by magnitude)
void zero_small_columns(Context* ctx, double* A, long long m, long long n, long long lda,
                        const double* vals, double rel);          // *out = sqrt(max(*in, 0))
// dot of two strided "vectors" with weights: out = |sum_{i,j..}|, used by BTRG finalize
// dst = flag ? a : b, flag = (*epsA > *epsB)  (device side choice, HOTRG projector pick)
void select_copy(Context* ctx, double* dst, const double* a, const double* b, long long n,
                 const double* eps_a, const double* eps_b, double* eps_out);

// ---- kernels: jacobi.cu ----
// One-sided Jacobi on the columns of G (m x n, ldg), accumulating V (n x n, ldv) when V != null.
// On return G = A*V has mutually orthogonal columns.  Returns number of sweeps.
int jacobi_orthogonalize(Context* ctx, double* G, long long m, long long n, long long ldg,
                         double* V, long long ldv);
// after orthogonalization: vals[j] = ||g_j|| (svd) or v_j . g_j (eigh, signed)
void column_values(Context* ctx, const double* G, long long m, long long n, long long ldg,
                   const double* V, long long ldv, double* vals, bool signed_rayleigh);
// rank[j] = position of |vals[j]| in descending order (stable); eps = sqrt(sum_{rank>=k} vals^2)
void rank_select(Context* ctx, const double* vals, long long n, long long k, int* rank,
                 double* dev_eps);
// dst[:, rank[j]] = src[:, j] * (normalize ? 1/|vals[j]| : 1) for rank[j] < k ; dst ld = ldd
void gather_columns(Context* ctx, const double* src, long long m, long long n, long long lds,
                    const int* rank, long long k, double* dst, long long ldd,
                    const double* vals, bool normalize);
void gather_values(Context* ctx, const double* vals, long long n, const int* rank, long long k,
                   double* out, bool absval);

}  // namespace tnr
