/* rsvdb.h -- C ABI of the B200-native randomized-SVD engine (librsvdb.so).
 *
 * This is the drop-in boundary for the rSVD hot path of AMSC22-23/rSVD_Kamaneh_Raganato_Terrana.  Each entry point
 * names the reference interface it replaces (file:line relative to the reference tree).  The C++ drop-in headers in
 * this directory (rSVD.hpp, SVD_class.hpp, QR.hpp, PM.hpp, matrixOperations.hpp) wrap these calls behind the
 * reference's own names; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every matrix is column-major FP64 (Eigen::MatrixXd layout, reference include/rSVD.hpp:9-10); `ld*` is the
 *     leading dimension in elements;
 *   - `*_host` entry points take host pointers and do the H2D / D2H copies themselves; `*_dev` entry points take
 *     device pointers that live on the context's GPU and enqueue on the context's stream WITHOUT synchronising;
 *   - every function returns RSVDB_OK (0) or a negative rsvdb_status; nothing throws across the boundary;
 *     rsvdb_last_error() returns a human-readable message for the last failure on that context;
 *   - there is no CPU fallback: without a CUDA device rsvdb_create fails with RSVDB_ERR_CUDA.
 */
#ifndef RSVDB_H
#define RSVDB_H

#include <stdint.h>

#if defined(__GNUC__)
#define RSVDB_API __attribute__((visibility("default")))
#else
#define RSVDB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsvdb_ctx rsvdb_ctx;

typedef enum {
  RSVDB_OK = 0,
  RSVDB_ERR_INVALID_ARGUMENT = -1, /* the reference throws std::invalid_argument (src/rSVD.cpp:122-123, src/matrixOperations.cpp:8-11) */
  RSVDB_ERR_CUDA = -2,
  RSVDB_ERR_NCCL = -3,
  RSVDB_ERR_ALLOC = -4,
  RSVDB_ERR_NO_CONVERGENCE = -5,
  RSVDB_ERR_UNSUPPORTED = -6
} rsvdb_status;

/* enum class SVDMethod { Jacobi, Power, ParallelJacobi } -- include/SVD_class.hpp:28-32 (same numeric values). */
typedef enum { RSVDB_SVD_JACOBI = 0, RSVDB_SVD_POWER = 1, RSVDB_SVD_PARALLEL_JACOBI = 2 } rsvdb_svd_method;

/* ---- context -------------------------------------------------------------------------------------------------- */
/* One context per process per GPU.  Owns a stream, scratch buffers and (optionally) an NCCL communicator. */
RSVDB_API int rsvdb_create(rsvdb_ctx** ctx, int device);
RSVDB_API int rsvdb_destroy(rsvdb_ctx* ctx);
/* Enqueue on a caller-owned cudaStream_t (e.g. torch's current stream).  NULL is a valid handle: the legacy default
 * stream.  rsvdb_use_own_stream() goes back to the context's private non-blocking stream. */
RSVDB_API int rsvdb_set_stream(rsvdb_ctx* ctx, void* cuda_stream);
RSVDB_API int rsvdb_use_own_stream(rsvdb_ctx* ctx);
RSVDB_API int rsvdb_synchronize(rsvdb_ctx* ctx);
RSVDB_API const char* rsvdb_last_error(const rsvdb_ctx* ctx);
RSVDB_API const char* rsvdb_version(void);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
RSVDB_API int64_t rsvdb_launch_count(const rsvdb_ctx* ctx);
/* Performance notes (process-wide counters, not errors) about the operands of *_dev products.
 * rsvdb_split_gemm_products: products whose big operand A had an odd leading dimension or a base pointer that is only
 *   8-byte aligned.  TMA cannot describe such a matrix with one tensor map; the engine describes its even and its odd columns
 *   separately (two maps, stride 2*lda) and runs the same FP64 tensor-core kernel -- same speed class, no copy of A.
 * rsvdb_generic_gemm_fallbacks: products that ran on the CUDA-core kernel instead (only a 1-column A that is also misaligned). */
RSVDB_API int64_t rsvdb_split_gemm_products(void);
RSVDB_API int64_t rsvdb_generic_gemm_fallbacks(void);
/* How the sketches of the rSVD path are orthonormalised (the reference: Eigen::HouseholderQR + thin Q, src/rSVD.cpp:60-68 and
 * the QR preconditioner of include/SVD_class.hpp:110-123).
 *   policy 0 (default): guarded CholeskyQR2 on the tensor cores -- two rounds of Gram matrix, Cholesky, triangular solve.  The
 *     engine MEASURES ||Q1^T Q1 - I|| after the first round; a Cholesky breakdown (rank-deficient sketch) or a value above 0.05
 *     (kappa(Y) beyond ~1e7) leaves the sketch untouched and the Householder TSQR runs instead, for that sketch and the rest of
 *     the factorisation.  Same basis to working precision (orthogonality ~1e-15), ~4x faster on well-conditioned sketches.
 *   policy 1: Householder TSQR always (also selected process-wide by the environment variable RSVDB_CHOLQR=0).
 * rsvdb_qr_dev (the QR class) is always Householder: its R carries the reference's sign convention.
 * rsvdb_qr_path_counts: sketches orthonormalised by either path since the context was created. */
RSVDB_API int rsvdb_set_qr_policy(rsvdb_ctx* ctx, int policy);
RSVDB_API int rsvdb_qr_path_counts(const rsvdb_ctx* ctx, int64_t* cholqr2, int64_t* householder);

/* Optional per-phase device timing (CUDA events on the context's stream).  rsvdb_phase_ms synchronises the stream,
 * writes the accumulated milliseconds per phase since the last call and clears them.
 * Phases: 0 A*X GEMMs, 1 A^T*Q GEMMs, 2 orthonormalisations (CholeskyQR2 / Householder TSQR), 3 small SVD (includes its own QR), 4 collectives, 5 other, 6 host<->device copies. */
#define RSVDB_NUM_PHASES 7
RSVDB_API int rsvdb_set_profiling(rsvdb_ctx* ctx, int enabled);
RSVDB_API int rsvdb_phase_ms(rsvdb_ctx* ctx, double* out_ms /* RSVDB_NUM_PHASES */);
/* Sweeps (negative: hit the cap without converging) and plane rotations of the last Jacobi small SVD. Synchronises. */
RSVDB_API int rsvdb_last_svd_info(rsvdb_ctx* ctx, int* sweeps, int* rotations);

/* ---- multi-GPU: one process per GPU, A row-sharded ------------------------------------------------------
 * The reference's only process boundary on this path is MPI_Gatherv/MPI_Bcast of Omega (src/rSVD.cpp:49,52) -- every MPI
 * rank repeats the whole computation.  Here rank p holds rows [offset_p, offset_p + m_p) of A; the engine all-reduces the
 * n x l partial sums of A^T Q and all-gathers the l x l TSQR R factors over NCCL (NVLink); everything else is shard-local.
 * The 128-byte NCCL unique id is produced on rank 0 and handed to every rank by the launcher (torch.distributed, MPI, ...). */
RSVDB_API int rsvdb_comm_unique_id(void* out_id_128_bytes);
RSVDB_API int rsvdb_comm_init(rsvdb_ctx* ctx, int nranks, int rank, const void* id_128_bytes);
RSVDB_API int rsvdb_comm_size(const rsvdb_ctx* ctx);
RSVDB_API int rsvdb_comm_rank(const rsvdb_ctx* ctx);

/* ---- dense building blocks, device pointers -------------------------------------------------------------------- */
/* Y (m x l) = A (m x n) * X (n x l).            Replaces Eigen `A * Omega`, `A * Q` at src/rSVD.cpp:59,66. */
RSVDB_API int rsvdb_gemm_an_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dX, int64_t ldx,
                                int l, double* dY, int64_t ldy);
/* Z (n x l) = A^T * Q, A m x n, Q m x l.        Replaces Eigen `A.transpose() * Q` at src/rSVD.cpp:63.
 * transpose_out != 0 stores B (l x n) = Q^T * A instead -- Eigen `Q.transpose() * A` at src/rSVD.cpp:89.
 * Shard-local: no reduction over ranks is performed here. */
RSVDB_API int rsvdb_gemm_at_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dQ, int64_t ldq,
                                int l, double* dZ, int64_t ldz, int transpose_out);
/* In-place thin Householder QR (TSQR): Y (rows x l) <- Q.  Replaces `HouseholderQR qr(Y); Q = qr.householderQ() *
 * Identity(rows, l)` at src/rSVD.cpp:60-61,64-65,67-68.  dR (optional, l x l, leading dimension l) receives R.
 * sharded != 0: Y is this rank's row block and the R factors are combined across ranks. */
RSVDB_API int rsvdb_qr_dev(rsvdb_ctx* ctx, double* dY, int64_t rows, int l, int64_t ldy, int sharded, double* dR);
/* The same step as the rSVD pipeline itself runs it: Y <- an orthonormal basis of its columns under the context's QR policy
 * (rsvdb_set_qr_policy: guarded CholeskyQR2, Householder TSQR when the guard refuses).  dR (optional, l x l, ld l): upper
 * triangular with Y_in = Q R; its diagonal is positive when CholeskyQR2 ran.  *path (optional): 0 CholeskyQR2, 1 Householder. */
RSVDB_API int rsvdb_orthonormalize_dev(rsvdb_ctx* ctx, double* dY, int64_t rows, int l, int64_t ldy, int sharded, double* dR, int* path);
/* intermediate_step(A, Q, Omega, l, q) -- include/rSVD.hpp:13, src/rSVD.cpp:57-70.  A: this rank's m_local x n row block. */
RSVDB_API int rsvdb_range_finder_dev(rsvdb_ctx* ctx, const double* dA, int64_t m_local, int64_t n, int64_t lda,
                                     const double* dOmega, int64_t ldo, int l, int q, double* dQ, int64_t ldq);
/* rSVD(A, U, S, V, l, method) -- include/rSVD.hpp:14, src/rSVD.cpp:72-133 -- with Omega and q (hard-coded 2 in the
 * reference, :83) as arguments.  U: m_local x k, S: k, V: n x k, k = min(l, n).  For RSVDB_SVD_POWER, U gets the
 * reference's identity-completed l columns and V holds the right singular vectors in COLUMNS (n x k); the drop-in
 * wrappers re-shape V to the reference's n x n rows layout (include/SVD_class.hpp:214). */
RSVDB_API int rsvdb_rsvd_dev(rsvdb_ctx* ctx, const double* dA, int64_t m_local, int64_t n, int64_t lda, const double* dOmega,
                             int64_t ldo, int l, int q, int method, uint64_t seed, double* dU, int64_t ldu, double* dS,
                             double* dV, int64_t ldv);
/* N(0,1) n x l test matrix on the device (counter-based generator; replaces generateOmega, src/rSVD.cpp:12-55, whose
 * std::random_device draw is not reproducible).  Every rank that passes the same seed gets the same Omega. */
RSVDB_API int rsvdb_generate_omega_dev(rsvdb_ctx* ctx, int64_t n, int l, uint64_t seed, double* dOmega, int64_t ldo);

/* ---- reference API mirrors, host pointers (H2D / D2H inside) ------------------------------------------------------ */
/* rSVD (src/rSVD.cpp:72-133).  Omega == NULL: drawn on the device from `seed`.  A is this rank's row block. */
RSVDB_API int rsvdb_rsvd_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                              uint64_t seed, int l, int q, int method, double* U, int64_t ldu, double* S, double* V, int64_t ldv);
/* intermediate_step (src/rSVD.cpp:57-70). */
RSVDB_API int rsvdb_intermediate_step_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega,
                                           int64_t ldo, int l, int q, double* Q, int64_t ldq);
/* generateOmega (src/rSVD.cpp:12-55). */
RSVDB_API int rsvdb_generate_omega_host(rsvdb_ctx* ctx, int64_t n, int l, uint64_t seed, double* Omega, int64_t ldo);
/* SVD<method>(A, r).compute() -- include/SVD_class.hpp:35-97.  k = min(m, n).
 *   Jacobi / ParallelJacobi: U m x k, S k, V n x k (columns), *found = k.
 *   Power: U m x m (identity-completed), S k, V n x dim (columns; dim = r ? r : k), *found = number of triplets before
 *   the reference's sigma < 1e-12 early exit (:198-209). */
RSVDB_API int rsvdb_svd_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, int method, int r, uint64_t seed,
                             double* U, int64_t ldu, double* S, double* V, int64_t ldv, int* found);
/* qr_decomposition_reduced / qr_decomposition_full (include/QR.hpp:15-16, src/QR.cpp:22-80) and the QR class
 * (image_compression/include/QR.hpp:13-64).  reduced (full == 0, m >= n): Q m x n, R n x n.  full: Q m x m, R m x n.
 * Householder-based; rows of R / columns of Q are sign-normalised so that diag(R) >= 0 (the Givens convention). */
RSVDB_API int rsvdb_qr_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, int full, double* Q, int64_t ldq,
                            double* R, int64_t ldr);
/* PM(A, B, sigma, u, v) -- include/PM.hpp:18, src/PM.cpp:4-81.  B = A^T A is not needed (see power.cu). */
RSVDB_API int rsvdb_pm_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, uint64_t seed, double* sigma,
                            double* u, double* v);
/* manualMatrixMultiply(A, B) -- include/matrixOperations.hpp:14, src/matrixOperations.cpp:7-28.
 * Returns RSVDB_ERR_INVALID_ARGUMENT when ka != kb (the reference throws std::invalid_argument, :8-11). */
RSVDB_API int rsvdb_gemm_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t ka, int64_t lda, const double* B, int64_t kb,
                              int64_t n, int64_t ldb, double* C, int64_t ldc);
/* ---- sparse inputs (CSR; int64 row pointers, int32 column indices, 0-based) ----------------------------------------
 * The reference densifies every MatrixMarket input before rSVD (tests/rSVD_test.cpp:54-57); these entry points run
 * the same algorithm (src/rSVD.cpp:57-133) on the CSR directly.  A is this rank's row block. */
RSVDB_API int rsvdb_rsvd_csr_host(rsvdb_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                                  const double* values, const double* Omega, int64_t ldo, uint64_t seed, int l, int q, int method,
                                  double* U, int64_t ldu, double* S, double* V, int64_t ldv);
RSVDB_API int rsvdb_rsvd_csr_dev(rsvdb_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* d_rowptr, const int32_t* d_colidx,
                                 const double* d_values, const double* dOmega, int64_t ldo, uint64_t seed, int l, int q, int method,
                                 double* dU, int64_t ldu, double* dS, double* dV, int64_t ldv);
/* Y (m x l, ROW-major) = A (CSR) * X (n x l, ROW-major): the SpMM building block (HBM-bound; see csrc/spmm.cu). */
RSVDB_API int rsvdb_csr_spmm_dev(rsvdb_ctx* ctx, int64_t m, const int64_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                 const double* dX_rowmajor, int l, double* dY_rowmajor);

/* ---- PCA front / back steps around the path ----------------------------------------------------
 * Reference: PCA/include/PCA_class.hpp.  data is m x n, rows = observations, columns = variables. */

/* PCA<method>(data, normalize)::initialize() (PCA_class.hpp:24-47): column means (:33), centring (:34), optional division
 * by the sample standard deviation (:38-41), then SVD<method> of the centred matrix (:45-46) -- one upload, everything
 * on the device.  Outputs as rsvdb_svd_host (same shapes per method); mean[n]; stddev[n] is written only when
 * normalize != 0 (may be NULL otherwise).  Fewer than 2 rows or 2 columns: RSVDB_ERR_INVALID_ARGUMENT with the
 * reference's message (assertDataValid, :50-54). */
RSVDB_API int rsvdb_pca_host(rsvdb_ctx* ctx, const double* data, int64_t m, int64_t n, int64_t ld, int normalize, int method, int r,
                             uint64_t seed, double* mean, double* stddev, double* U, int64_t ldu, double* S, double* V,
                             int64_t ldv, int* found);

/* Column statistics of device data (row shard): mean[j] (PCA_class.hpp:33) and, when d_stddev != NULL, the sample
 * standard deviation of the centred column (:39).  Sums are all-reduced over the ranks of rsvdb_comm_init. */
RSVDB_API int rsvdb_column_stats_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, double* d_mean,
                                     double* d_stddev);
/* In place: A(i,j) <- (A(i,j) - mean[j]) / (d_stddev ? d_stddev[j] : 1)   (PCA_class.hpp:34,40). */
RSVDB_API int rsvdb_center_columns_dev(rsvdb_ctx* ctx, double* dA, int64_t m, int64_t n, int64_t lda, const double* d_mean,
                                       const double* d_stddev);

/* Randomized PCA: rSVD (src/rSVD.cpp:72-133) of the centred (and scaled) matrix WITHOUT materialising it -- the six
 * passes stream the original A and the centring enters as rank-1 corrections of the skinny operands
 * ((A - 1 mu^T) D X = A (D X) - 1 (mu^T D X)).  d_stddev == NULL: centring only.  Same outputs as rsvdb_rsvd_dev. */
RSVDB_API int rsvdb_rpca_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, const double* d_mean,
                             const double* d_stddev, const double* dOmega, int64_t ldo, int l, int q, int method, double* dU,
                             int64_t ldu, double* dS, double* dV, int64_t ldv);
/* Host mirror: upload, column statistics, rsvdb_rpca_dev, download.  Omega == NULL: drawn on the device from seed. */
RSVDB_API int rsvdb_rpca_host(rsvdb_ctx* ctx, const double* A, int64_t m, int64_t n, int64_t lda, int normalize, const double* Omega,
                              int64_t ldo, uint64_t seed, int l, int q, int method, double* mean, double* stddev, double* U,
                              int64_t ldu, double* S, double* V, int64_t ldv);

/* projectToPCA (PCA_class.hpp:93-95): out (r x k) = (data - 1 mean^T) * V,  data r x n, V n x k. */
RSVDB_API int rsvdb_pca_project_host(rsvdb_ctx* ctx, const double* data, int64_t r, int64_t n, int64_t ld, const double* mean,
                                     const double* V, int64_t ldv, int k, double* out, int64_t ldout);
/* reconstructFromPCA (PCA_class.hpp:97-99): out (r x n) = pc * V^T + 1 mean^T,  pc r x k, V n x k. */
RSVDB_API int rsvdb_pca_reconstruct_host(rsvdb_ctx* ctx, const double* pc, int64_t r, int k, int64_t ldp, const double* mean,
                                         const double* V, int64_t ldv, int64_t n, double* out, int64_t ldout);

/* ---- POD wrappers --------------------------------------------------------------------------------
 * Reference: POD/ParametricDiffusion1D/src/POD.cpp.  S is the Nh x ns snapshot matrix.
 * variant: 0 naive_POD (:116-134), 1 standard_POD (:136-224), 2 energy_POD (:226-336, needs Xh Nh x Nh),
 *          3 weight_POD (:338-461, needs Xh and D ns x ns).
 * svd_type (perform_SVD, :42-114): 0 SVD<Power>(A, r), 1 SVD<Jacobi>, 2 SVD<ParallelJacobi>, 3/4/5 rSVD(A, ..., r, Power /
 *          Jacobi / ParallelJacobi); anything else: RSVDB_ERR_INVALID_ARGUMENT with the reference's message (it exits).
 * Omega (optional): the sketch for svd_type 3-5, (columns of the matrix handed to rSVD: ns for the naive variant, else
 *          min(ns, Nh)) x r; NULL: drawn on the device from seed.
 * Outputs: W (Nh x *N, the first *N of w_cols_full columns are written), sigma (sigma_len values: the singular values of
 * the CORRELATION matrix for variants 1-3, exactly what the reference stores), *N = modes kept by the energy criterion
 * (:203-219).  rsvdb_pod_shape gives w_cols_full / sigma_len for sizing the buffers. */
RSVDB_API int rsvdb_pod_shape(int variant, int64_t Nh, int64_t ns, int r, int svd_type, int64_t* w_cols_full, int64_t* sigma_len);
RSVDB_API int rsvdb_pod_host(rsvdb_ctx* ctx, int variant, const double* S, int64_t Nh, int64_t ns, int64_t lds, const double* Xh,
                             int64_t ldx, const double* D, int64_t ldd, int r, double tol, int svd_type, uint64_t seed,
                             const double* Omega, int64_t ldo, double* W, int64_t ldw, double* sigma, int* N);
/* Device-pointer twin: W must hold Nh x w_cols_full, all of which are written; synchronises the stream once. */
RSVDB_API int rsvdb_pod_dev(rsvdb_ctx* ctx, int variant, const double* dS, int64_t Nh, int64_t ns, int64_t lds, const double* dXh,
                            int64_t ldx, const double* dD, int64_t ldd, int r, double tol, int svd_type, uint64_t seed,
                            const double* dOmega, int64_t ldo, double* dW, int64_t ldw, double* d_sigma, int* N);

/* ---- Image::compress-shaped driver ----------------------------------------------------------------
 * Reference: image_compression/src/image_com.cpp.  image is m x n column-major (Image::image_matrix). */

/* Image::normalize() (:251-264) + Image::compress(k) (:288-317): min-max normalisation to [0,1] when normalize != 0 (skipped
 * with a warning in the reference when min >= max; here *original_min / *original_max report the range either way), then the
 * older-API rSVD(A, U, S, V, l) (image_compression/src/rSVD.cpp:77-118: q = 1, power-method back-end) with l = k + 10
 * (k = -1: min(m, n) / 4).  Outputs U m x l, S l, V n x l (columns), *degree = l.  Omega (n x l) optional. */
RSVDB_API int rsvdb_image_compress_host(rsvdb_ctx* ctx, const double* image, int64_t m, int64_t n, int64_t ld, int k, int normalize,
                                        const double* Omega, int64_t ldo, uint64_t seed, double* original_min, double* original_max,
                                        double* U, int64_t ldu, double* S, double* V, int64_t ldv, int* degree);
/* Image::normalize() (inverse == 0, :251-264: finds min / max, writes them to *original_min / *original_max and maps the
 * image to [0,1] in place) or Image::deNormalize() (inverse != 0, :270-281: uses the given range).  A no-op on the data when
 * min >= max (the reference prints a warning); returns RSVDB_OK either way. */
RSVDB_API int rsvdb_image_normalize_host(rsvdb_ctx* ctx, double* image, int64_t m, int64_t n, int64_t ld, int inverse,
                                         double* original_min, double* original_max);
/* Image::reconstruct() (:184-190) + Image::deNormalize() (:270-281): out (m x n) = U diag(S) V^T, then x * (max - min) + min when
 * denormalize != 0 and min < max -- the affine map is applied in the same pass over the result. */
RSVDB_API int rsvdb_image_reconstruct_host(rsvdb_ctx* ctx, const double* U, int64_t m, int64_t ldu, const double* S, const double* V,
                                           int64_t n, int64_t ldv, int l, int denormalize, double original_min, double original_max,
                                           double* out, int64_t ldout);

/* Number of power iterations PM runs for an n-column matrix (src/PM.cpp:25-28). */
RSVDB_API int rsvdb_pm_iterations(int64_t ncols);

#ifdef __cplusplus
}
#endif
#endif /* RSVDB_H */
