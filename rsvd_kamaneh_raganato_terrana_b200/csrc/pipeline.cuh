// Device-side orchestration of the rSVD path (see pipeline.cu).
#pragma once
#include <cstdint>
#include "context.cuh"
#include "pca.cuh"

namespace rsvdb {

// smallest even leading dimension >= rows (TMA needs 16-byte column strides)
inline int64_t even_ld(int64_t rows) { return rows < 2 ? 2 : ((rows + 1) & ~int64_t(1)); }

// Orthonormalise the columns of Y (rows x l, ldy) in place with Householder TSQR.  sharded: Y is this rank's row block of
// a panel distributed over c->nranks ranks.  *R (optional) receives a device pointer to the l x l upper-triangular factor
// (leading dimension l), valid until the next QR on this context.
int qr_inplace(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool sharded, const double** R);

// Orthonormal basis of the columns of Y in place, as the pipeline needs it (cholqr.cu): guarded CholeskyQR2 when the sketch is
// numerically full-rank and well enough conditioned (decided from the measured ||Q1^T Q1 - I||), else qr_inplace.  *R (optional):
// l x l upper triangular with Y_in = Q R (its diagonal is positive on the fast path, Householder-signed otherwise).
int orthonormalize(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool sharded, const double** R);

// A still lives in host memory when the path starts (the host-pointer entry points): the upload is cut into row blocks
// on the side stream and the first product Y = A * Omega consumes each block as it lands, so all but the last block's
// share of that pass hides under the PCIe transfer.
struct HostUpload {
  const double* A = nullptr;   // host matrix, column-major
  int64_t lda = 0;
};

// intermediate_step (reference src/rSVD.cpp:57-70): Q <- range finder with q power iterations.  A is this rank's row
// block (m_local x n).  Q is m_local x l.  up != nullptr: the device buffer A is filled from up->A on the way.
int range_finder(rsvdb_ctx* c, const double* A, int64_t m_local, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                 int l, int q, double* Q, int64_t ldq, const HostUpload* up = nullptr, const Centering* cen = nullptr);

// SVD<Jacobi|ParallelJacobi> (include/SVD_class.hpp:101-180, :224-333) of a device matrix M (r x c, ldm), or of its
// transpose when Mt != nullptr is given instead (c x r, ldmt).  U r x k, S k, V c x k, k = min(r,c).
int small_svd_jacobi(rsvdb_ctx* c, const double* M, int64_t ldm, const double* Mt, int64_t ldmt, int64_t r, int64_t cdim,
                     double* U, int64_t ldu, double* S, double* V, int64_t ldv);

// rSVD (src/rSVD.cpp:72-133) on device data; A is this rank's row block.  U m_local x k, S k, V n x k, k = min(l, n).
// cen != nullptr: the factorisation is that of (A - 1 mu^T) diag(inv_sd) -- the matrix PCA_class.hpp:30-41 would
// materialise -- computed from the uncentred A through rank-1 corrections (pca.cuh).
int rsvd_device(rsvdb_ctx* c, const double* A, int64_t m_local, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                int l, int q, int method, double* U, int64_t ldu, double* S, double* V, int64_t ldv, uint64_t seed,
                const HostUpload* up = nullptr, const Centering* cen = nullptr);

// SVD<Power> (include/SVD_class.hpp:184-219 + src/PM.cpp) -- power.cu.  Mt is the TRANSPOSE (c x r) of the data matrix
// and is deflated in place.  U r x r (identity-completed), S min(r,c), V c x dim (columns = right singular vectors).
int small_svd_power_t(rsvdb_ctx* c, double* Mt, int64_t ldmt, int64_t r, int64_t cdim, int rdim, uint64_t seed,
                      double* U, int64_t ldu, int u_cols, double* S, double* V, int64_t ldv, int* found_host);
int pm_iterations(int64_t ncols);
int copy2d(rsvdb_ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int cols);
int transpose2d(rsvdb_ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols);

}  // namespace rsvdb
