// PCA front/back steps around the rSVD path (reference PCA/include/PCA_class.hpp): column statistics, centring /
// scaling, and the rank-1 corrections that let the GEMM passes run on the UNCENTRED matrix.
#pragma once
#include <cstdint>
#include "context.cuh"

namespace rsvdb {

// Implicit centring  Ac = (A - 1 mu^T) * diag(inv_sd)  applied around the products of the range finder:
//   Ac * X   = A * (D X) - 1 * (mu^T D X)                      (one n x l row scaling, one l-vector, one m x l update)
//   Ac^T * Q = D * (A^T Q - mu * (1^T Q))                      (one l-vector of column sums, one n x l update)
// mu / inv_sd are device vectors of length n; inv_sd == nullptr means D = I.
struct Centering {
  const double* mu = nullptr;
  const double* inv_sd = nullptr;
};

// mean[j] = sum_i A(i,j) / M and, when stddev != nullptr, stddev[j] = sqrt(sum_i (A(i,j)-mean[j])^2 / (M-1)) with M the
// GLOBAL row count (sums are all-reduced over the row shards).  PCA_class.hpp:33,39.  inv_sd (optional) = 1 / stddev.
int column_stats(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, double* mean, double* stddev, double* inv_sd);
// A(i,j) <- (A(i,j) - mean[j]) / (sd ? sd[j] : 1)     PCA_class.hpp:34,40
int center_columns(rsvdb_ctx* c, double* A, int64_t m, int64_t n, int64_t lda, const double* mean, const double* sd);

// pca_ws layout (doubles): [column_stats scratch | inv_sd: n | Xs: n x l | w: l | s: l]
struct PcaScratch {
  static size_t pad(size_t d) { return (d + 31) & ~size_t(31); }
  static size_t inv_sd_off(int64_t n) { return pad(2 * (size_t)n + 8); }
  static size_t xs_off(int64_t n) { return inv_sd_off(n) + pad((size_t)n); }
  static size_t total(int64_t n, int l) { return xs_off(n) + pad((size_t)n * l) + 2 * pad((size_t)l) + 64; }
};

// pieces of the implicit-centring products (see Centering)
int scale_rows_copy(rsvdb_ctx* c, const double* X, int64_t ldx, double* Xs, int64_t lds, int64_t n, int l, const double* inv_sd);
int weighted_colsum(rsvdb_ctx* c, const double* X, int64_t ld, int64_t rows, int cols, const double* wgt, double* out);
int sub_col_const(rsvdb_ctx* c, double* Y, int64_t ld, int64_t rows, int cols, const double* w);
int rank1_correct(rsvdb_ctx* c, double* Z, int64_t ld, int64_t rows, int cols, const double* mu, const double* s, const double* inv_sd);
// out(i,j) += sign * mean[j]     (projectToPCA / reconstructFromPCA epilogues, PCA_class.hpp:93-100)
int add_row_vector(rsvdb_ctx* c, double* out, int64_t ld, int64_t rows, int64_t cols, const double* mean, double sign);

}  // namespace rsvdb
