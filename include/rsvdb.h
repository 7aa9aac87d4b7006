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

/* ---- dense building blocks, device pointers -------------------------------------------------------------------- */
/* Y (m x l) = A (m x n) * X (n x l).            Replaces Eigen `A * Omega`, `A * Q` at src/rSVD.cpp:59,66. */
RSVDB_API int rsvdb_gemm_an_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dX, int64_t ldx,
                      int l, double* dY, int64_t ldy);
/* Z (n x l) = A^T * Q, A m x n, Q m x l.        Replaces Eigen `A.transpose() * Q` at src/rSVD.cpp:63.
 * transpose_out != 0 stores B (l x n) = Q^T * A instead -- Eigen `Q.transpose() * A` at src/rSVD.cpp:89. */
RSVDB_API int rsvdb_gemm_at_dev(rsvdb_ctx* ctx, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dQ, int64_t ldq,
                      int l, double* dZ, int64_t ldz, int transpose_out);

#ifdef __cplusplus
}
#endif
#endif /* RSVDB_H */
