/* libtnrcuda -- C ABI of the B200 (sm_100a) engine behind TNRKit's coarse-graining
 * hot path.  Plain pointers and sizes only; every tensor is FP64, column major
 * (first index fastest), i.e. exactly the memory layout of the dense block of a
 * TensorKit `TensorMap{Float64, ComplexSpace}` whose legs are (codomain..., domain...).
 * Device pointers are raw CUDA device addresses owned by the caller (the Julia
 * host's device-resident storage type) unless allocated with tnr_malloc.
 *
 * Every entry point cites the reference interface it replaces
 * (paths relative to VictorVanthilt/TNRKit.jl v0.5.1).
 *
 * All functions return 0 on success; nonzero on error (1 = invalid argument,
 * 2 = CUDA error, 3 = internal).  tnr_last_error(ctx) returns the message.
 * There is no CPU fallback anywhere: without a CUDA device tnr_create fails.
 */
#ifndef TNRCUDA_H
#define TNRCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnr_context tnr_context;

/* scheme kinds for tnr_step_out_dims */
enum {
    TNR_TRG = 0,
    TNR_BTRG = 1,
    TNR_HOTRG = 2,
    TNR_ATRG = 3,
    TNR_HOTRG_3D = 4,
    TNR_ATRG_3D = 5
};

/* ---- context ----------------------------------------------------------- */
int tnr_version(void);
/* `stream` is a cudaStream_t; NULL selects the legacy default stream.  All work of the
 * context is enqueued on that stream, so it is ordered with the caller's own work there. */
int tnr_create(int device, void* stream, tnr_context** out);
int tnr_destroy(tnr_context* ctx);
const char* tnr_last_error(tnr_context* ctx);
int tnr_synchronize(tnr_context* ctx);
/* counters of this library's own kernel launches (bench.py "gpu_launches") */
int tnr_get_counters(tnr_context* ctx, uint64_t* launches, uint64_t* gemm_launches,
                     double* gemm_flops, double* permute_bytes);
int tnr_reset_counters(tnr_context* ctx);
/* counters of the warp-specialised TMA GEMM (subset of gemm_launches) */
int tnr_get_tma_launches(tnr_context* ctx, uint64_t* tma_gemm_launches);
/* any counter by name: "launches", "gemm_launches", "grouped_gemm_launches",
 * "tma_gemm_launches", "gemm_flops", "permute_bytes", "permute_bulk_launches", ... */
int tnr_get_counter(tnr_context* ctx, const char* name, double* value);
/* engine options: "disable_tma" = 1 forces the cp.async GEMM for every layout (A/B tests);
 * "disable_subspace", "disable_block_jacobi", "disable_precondition" switch the fast SVD/eigh
 * paths off; "disable_qr" = 1 replaces the Householder QR stage of tall problems by the round-1
 * Gram-preconditioned Jacobi; "ozaki" = S (0 = off, default) enables the INT8 emulation engine with S planes;
 * "ozaki_crt" = N (0 = off, default; 14..18) selects its CRT variant with N moduli instead;
 * "permute_bulk" = 1 (default) | 0: the TMA-fed tiled copy kernel (cp.async.bulk reads) for every
 * strided copy whose source pieces are 16-byte aligned; "permute_unroll" = 1 | 2 | 4 (default) and
 * "permute_tile" = 32 | 48 | 64 | 96 (default) select variants of the fallback permute kernels.
 * "permute_tpc" (1..8) and "permute_chunk_below" (bytes) tune the TMA-fed copy;
 * "hotrg3d_pk_budget_mb": megabytes of absorbed operands Pk_d the HOTRG_3D z-compression holds
 * at once (default 49152; the d loop is blocked into windows, also capped by the free memory).
 * "disable_cholqr" = 1: the subspace solvers re-orthonormalise their bases by Householder QR
 * instead of CholeskyQR2; "jacobi_max_bc" = 4 | 8 (default) | 16: largest column block of the
 * shared-memory Jacobi kernels (A/B tests).
 * Unknown keys and out-of-range values are errors. */
int tnr_set_option(tnr_context* ctx, const char* key, int64_t value);
/* CUDA-event timing of the dominant kernel (DMMA GEMM launches above 1e11 flop) on the
 * library stream; read returns the summed milliseconds, flops and the launch count. */
int tnr_gemm_timing(tnr_context* ctx, int enable);
int tnr_gemm_timing_read(tnr_context* ctx, double* ms_total, double* flops_total, int64_t* count);

/* ---- device memory (replaces the host Array storage of TensorMap.data) -- */
int tnr_malloc(tnr_context* ctx, size_t bytes, void** dptr);
int tnr_free(tnr_context* ctx, void* dptr);
int tnr_upload(tnr_context* ctx, void* dst_dev, const void* src_host, size_t bytes);
int tnr_download(tnr_context* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- primitives --------------------------------------------------------- */
/* C = alpha op(A) op(B) + beta C  (BLAS dgemm semantics; replaces LinearAlgebra.BLAS.gemm!
 * as called by TensorOperations.tensorcontract! for every `@tensor` line of the step! bodies,
 * e.g. src/schemes/trg.jl:42, src/schemes/hotrg3d.jl:116-120).  FP64 tensor cores (DMMA). */
int tnr_gemm(tnr_context* ctx, char transa, char transb, int m, int n, int k, double alpha,
             const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
             int64_t ldc);
int tnr_gemm_strided_batched(tnr_context* ctx, char transa, char transb, int m, int n, int k,
                             double alpha, const double* A, int64_t lda, int64_t strideA,
                             const double* B, int64_t ldb, int64_t strideB, double beta,
                             double* C, int64_t ldc, int64_t strideC, int batch);
/* C(m x n) = A^T B with A: k x m, B: k x n (column major; the 'T','N' case of tnr_gemm, alpha = 1,
 * beta = 0) computed on the INT8 tensor cores by error-free splitting (Ozaki scheme, tcgen05 +
 * TMEM): FP64-level accuracy relative to (row max) x (column max), ~2x the DMMA rate.  Opt-in:
 * tnr_set_option(ctx, "ozaki", 8) selects 8 digit planes (36 exact INT8 products) and also routes
 * the chi^3 x chi^3 x chi^3 chunk contraction of tnr_hotrg3d_step / _substep through it. */
int tnr_gemm_ozaki(tnr_context* ctx, int m, int n, int k, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double* C, int64_t ldc);
/* Grouped GEMM: `count` independent problems C_g = alpha op(A_g) op(B_g) + beta C_g in ONE
 * launch -- the per-coupled-sector block products of a Z2 / ZN / U1 block-sparse contraction
 * (TensorKit `mul!` loops over `blocks(t)`; e.g. the per-sector products behind every
 * `@tensor` line of src/schemes/trg.jl:42 and btrg.jl:86-94 for `Z2Irrep` / `ZNIrrep` tensors). */
typedef struct {
    int32_t m, n, k;
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    double* C;
    int64_t ldc;
} tnr_gemm_problem;
int tnr_gemm_grouped(tnr_context* ctx, char transa, char transb, int count,
                     const tnr_gemm_problem* problems, double alpha, double beta);
/* dst leg k = src leg perm[k] (0-based); replaces TensorKit.permute / transpose
 * (src/schemes/hotrg3d.jl:134, atrg.jl:40, atrg3d.jl:37, trg.jl:40). */
int tnr_permute(tnr_context* ctx, const double* src, double* dst, int rank, const int64_t* dims,
                const int* perm);
/* dst[sum_j i_j*dstride_j] = src[sum_j i_j*sstride_j] for all 0 <= i_j < dims[j]: block
 * gather/scatter between the tuple blocks and the coupled-sector matrices of an abelian
 * symmetric tensor (TensorKit's fusion-tree block <-> matrix-block views, `permute` on
 * `Z2Irrep` / `ZNIrrep` tensors). */
int tnr_strided_copy(tnr_context* ctx, const double* src, double* dst, int rank,
                     const int64_t* dims, const int64_t* sstride, const int64_t* dstride);
/* sum over the multi-index of w_0[i_0]..w_r[i_r] * src[sum_j i_j*stride_j] (weights may be
 * NULL): partial traces `@tensor T[1 2; 2 1]` of finalize! on one block
 * (src/utility/finalize.jl:4-14).  rank <= 4.  Result returned to the host. */
int tnr_strided_sum(tnr_context* ctx, const double* src, int rank, const int64_t* dims,
                    const int64_t* stride, const double* const* weights, double* sum_out);
/* x *= alpha  (`scheme.T /= n`, finalize.jl:6) */
int tnr_scale(tnr_context* ctx, double* x, int64_t n, double alpha);
/* A(m x n, lda): rows ? A[i,:] *= f(s[i]) : A[:,j] *= f(s[j]); f = identity (mode 0),
 * sqrt (mode 1: `U * sqrt(S)`, projectors.jl:218) or pseudopow(.,p) (mode 2: btrg.jl:51-60) */
int tnr_diag_scale(tnr_context* ctx, double* A, int64_t m, int64_t n, int64_t lda, const double* s,
                   int rows, int mode, double p);
/* A viewed as [m1][n][m2] (column major): A[i,j,k] *= f(s[j]) -- a diagonal bond tensor applied
 * to any leg (the S1 / S2 factors of the BTRG contraction, btrg.jl:86-94) */
int tnr_axis_scale(tnr_context* ctx, double* A, int64_t m1, int64_t n, int64_t m2, const double* s,
                   int mode, double p);
/* out[i] = f(s[i]) with the same modes (S_b = pseudopow(S, k), btrg.jl:66) */
int tnr_vec_map(tnr_context* ctx, const double* s, double* out, int64_t n, int mode, double p);
/* Sector-global truncrank on the device: rank[j] = position of |vals[j]| in descending order
 * over the concatenated spectra of all sectors, eps = 2-norm of the values of rank >= k.
 * Only the n ranks travel to the host (block shapes are host metadata). */
int tnr_topk_select(tnr_context* ctx, const double* vals, int64_t n, int64_t k, int32_t* rank_host,
                    double* eps_out);
/* Pairwise contraction by single-character leg labels (einsum for two operands; labels that
 * appear in A and B but not in C are summed).  Replaces one binary `@tensor` contraction
 * (TensorOperations.tensorcontract!).  C is written compact in label order labelsC. */
int tnr_contract(tnr_context* ctx, const double* A, int rankA, const int64_t* dimsA,
                 const char* labelsA, const double* B, int rankB, const int64_t* dimsB,
                 const char* labelsB, double* C, const char* labelsC);
/* svd_trunc(T; trunc = truncrank(chi)) with the first `ncod` legs as codomain
 * (MatrixAlgebraKit via TensorKit; src/utility/projectors.jl:213-219, btrg.jl:63, atrg.jl:38).
 * k = min(chi, min(rows, cols)).  U: rows x k, S: k, Vt: k x cols, eps: 2-norm of discarded. */
int tnr_svd_trunc(tnr_context* ctx, const double* T, int rank, const int64_t* dims, int ncod,
                  int chi, double* U, double* S, double* Vt, int64_t* k_out, double* eps_out);
/* Thin QR of a tall column-major matrix A (m x n, m >= n): blocked Householder (panels of 32
 * columns, compact WY, trailing update as FP64 tensor-core GEMMs) -- the factorization behind
 * TensorKit `left_orth` / `right_orth` (src/schemes/atrg3d.jl:53-56) and the first stage of the
 * truncated SVD of tall matrices (QR, then one-sided Jacobi of R in shared memory, then the
 * top-chi selection).  Q (m x n, orthonormal columns) and / or R (n x n upper triangular, LAPACK
 * sign convention) may be NULL. */
int tnr_qr(tnr_context* ctx, const double* A, int64_t m, int64_t n, double* Q, double* R);
/* The factor R of `_, R = left_orth(T)` (src/schemes/atrg3d.jl:53-56) for a tall matricization
 * (rows = first `ncod` legs >= cols): the upper-triangular Householder R (cols x cols,
 * R^T R = T^T T); with the option "disable_qr" the round-1 form R = Sigma V^T, which differs by an
 * orthogonal gauge on the new bond that cancels in the projectors built from it (atrg3d.jl:58-66).
 * The isometry Q is never formed.  Used chunk by chunk (TSQR) by the factored ATRG_3D step. */
int tnr_orth_r(tnr_context* ctx, const double* T, int rank, const int64_t* dims, int ncod,
               double* R);
/* A factor L (n x n, column major; columns >= *rank_out are zero) with L L^T = G for a symmetric
 * positive semidefinite G (n x n, not modified): diagonally pivoted Cholesky without row
 * exchanges, stopped at the numerical rank (largest remaining diagonal entry <= 8 eps max_i G_ii).
 * With G = A^T A, L^T is an R factor of A up to a left orthogonal gauge -- the only property of
 * `_, R = left_orth(...)` / `R, _ = right_orth(...)` that src/schemes/atrg3d.jl:53-66 uses; the
 * factored ATRG_3D step obtains its four R factors this way from the chi^2 x chi^2 Gram matrices of
 * the matricizations of YD / AX (one launch per column + one DMMA GEMM per 64 columns). */
int tnr_psd_factor(tnr_context* ctx, const double* G, int64_t n, double* L, int64_t* rank_out);
/* x[i] = uniform(-1, 1) from the splitmix64 hash of (seed, i): the start block of the subspace
 * iterations, generated where it is used (deterministic, identical on every rank of a sharded run). */
int tnr_fill_random(tnr_context* ctx, double* x, int64_t n, uint64_t seed);
/* In place: A (m x n, column major, m >= n) <- Q with orthonormal columns and the same column
 * space (Q = A R^-1, R the Cholesky factor of A^T A; CholeskyQR2: Gram matrix and A R^-1 on the FP64
 * tensor cores, twice).  The basis refresh BETWEEN the Rayleigh-Ritz steps of the block subspace
 * iteration that serves `svd_trunc(...; trunc = truncrank(chi))` on operators too large to
 * decompose (src/schemes/atrg3d.jl:35,43 at chi = 48); the Rayleigh-Ritz steps themselves, which
 * decide the result, use tnr_svd_trunc.  *refused_out = 1 and A untouched when A^T A is not
 * safely positive definite (rank deficient, cond(A) >~ 1e5) or n > 152: take tnr_qr /
 * tnr_svd_trunc then. */
int tnr_orthonormalize(tnr_context* ctx, double* A, int64_t m, int64_t n, int* refused_out);
/* eigh_trunc!(project_hermitian!(MM); trunc = truncrank(chi)) (src/schemes/hotrg.jl:106,114,
 * hotrg3d.jl:94-98).  Keeps the chi eigenvalues of largest magnitude.  MM is n x n and is
 * not modified.  W: k signed eigenvalues, V: n x k. */
int tnr_eigh_trunc(tnr_context* ctx, const double* MM, int64_t n, int chi, double* W, double* V,
                   int64_t* k_out, double* eps_out);

/* ---- scheme step! / finalize! bodies ------------------------------------ */
/* Output leg dimensions of one step! for a tensor with legs `dims` (4 for 2D, 6 for 3D). */
int tnr_step_out_dims(int scheme, const int64_t* dims, int chi, int64_t* dims_out);

/* step!(::TRG, truncrank(chi))      src/schemes/trg.jl:38-44 */
int tnr_trg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                 int64_t* dims_out);
/* step!(::BTRG, truncrank(chi))     src/schemes/btrg.jl:62-97.  S1/S2 are the DIAGONALS of the
 * bond tensors (they are identity / DiagonalTensorMap-valued in the reference), lengths
 * dims[1] and dims[0]; outputs have lengths dims_out[1], dims_out[0]. */
int tnr_btrg_step(tnr_context* ctx, const double* T, const int64_t* dims, const double* S1,
                  const double* S2, double k, int chi, double* Tout, int64_t* dims_out,
                  double* S1out, double* S2out);
/* step!(::HOTRG, trunc)             src/schemes/hotrg.jl:155-161 */
int tnr_hotrg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                   int64_t* dims_out);
/* step!(::ATRG, trunc)              src/schemes/atrg.jl:37-45 */
int tnr_atrg_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                  int64_t* dims_out);
/* step!(::HOTRG_3D, trunc)          src/schemes/hotrg3d.jl:131-139 (three _step! + permutes) */
int tnr_hotrg3d_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                     double* Tout, int64_t* dims_out);
/* _step!(::HOTRG_3D, trunc)         src/schemes/hotrg3d.jl:124-129, restricted to the slices
 * Tout[:, :, :, :, :, f] with f_begin <= f < f_end of the new open x-bond (multi-GPU sharding:
 * each rank computes its own slices, the host all-gathers along the last leg).  Tout must be
 * the full-size buffer; the permute of step! is NOT applied (call tnr_permute). */
int tnr_hotrg3d_substep(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                        double* Tout, int64_t* dims_out, int64_t f_begin, int64_t f_end);
/* Same z-compression, with the all-gather fused into the producing kernel: `Tout_peers[r]` is
 * rank r's full-size T' buffer as a peer-mapped device pointer (NVLink / NVSwitch; e.g. from
 * torch symmetric memory or cudaIpcOpenMemHandle), `self` this rank's index.  Every chi^4 slab
 * T'[:, :, :, d, :, f] is stored to ALL buffers by the kernel that produces it, so the transfer
 * overlaps the chunk loop and no collective follows; the caller only needs a cross-rank barrier
 * before reading its buffer. */
int tnr_hotrg3d_substep_peers(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                              double* const* Tout_peers, int npeers, int self, int64_t* dims_out,
                              int64_t f_begin, int64_t f_end);
/* The two halves of _step!(::HOTRG_3D) as separate entries, so that the N ranks of a sharded run
 * do not all repeat the projector work (the replicated part that limited scaling in round 1).
 * tnr_hotrg3d_proj_half: ONE of the four truncated eigendecompositions of
 * _get_hotrg3d_xproj / _yproj (src/schemes/hotrg3d.jl:83-108): which = 0: x-bond, MM^dagger
 * (_get_MMdag_3d, :47-66); 1: x-bond, M^dagger M (_get_MdagM_3d, :68-87); 2 / 3: the same for the
 * y-bond (after the permutation of :103-108).  Writes U (n x k, column major, n = D^2,
 * k = min(chi, n)) followed by the truncation error eps (one double) to `out` (n*k + 1 doubles,
 * device memory).  Ranks compute different halves and exchange the packed results (broadcast).
 * tnr_hotrg3d_contract: hotrg3d.jl:116-120 for the slices f_begin <= f < f_end with the four
 * packed halves given back to back in `halves` (x-left, x-right, y-left, y-right); the choice
 * `(eps > eps') ? U' : U` of hotrg3d.jl:96 is made on the device.  `Tout_peers` / `npeers` /
 * `self` as in tnr_hotrg3d_substep_peers (npeers = 1: a single local buffer). */
int tnr_hotrg3d_proj_half(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                          int which, double* out);
int tnr_hotrg3d_contract(tnr_context* ctx, const double* T, const int64_t* dims, int chi,
                         const double* halves, double* const* Tout_peers, int npeers, int self,
                         int64_t* dims_out, int64_t f_begin, int64_t f_end);
/* step!(::ATRG_3D, trunc)           src/schemes/atrg3d.jl:85-97 */
int tnr_atrg3d_step(tnr_context* ctx, const double* T, const int64_t* dims, int chi, double* Tout,
                    int64_t* dims_out);

/* finalize!(::Union{TRG,ATRG,HOTRG})  src/utility/finalize.jl:4-8 : n = |T[1 2;2 1]|, T /= n */
int tnr_finalize_2d(tnr_context* ctx, double* T, const int64_t* dims, double* norm_out);
/* finalize!(::BTRG)                   src/utility/finalize.jl:10-14 */
int tnr_finalize_btrg(tnr_context* ctx, double* T, const int64_t* dims, const double* S1,
                      const double* S2, double* norm_out);
/* finalize!(::HOTRG_3D / ::ATRG_3D)   src/utility/finalize.jl:56-66 : n = |T[1 1;2 3 2 3]| */
int tnr_finalize_3d(tnr_context* ctx, double* T, const int64_t* dims, double* norm_out);

#ifdef __cplusplus
}
#endif
#endif /* TNRCUDA_H */
