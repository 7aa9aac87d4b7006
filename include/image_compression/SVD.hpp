// Drop-in for reference image_compression/include/SVD.hpp (image_compression/src/SVD.cpp:3-55):
//   singularValueDecomposition(A, sigma, U, V, dim)  -- `dim` dominant triplets by power iteration + rank-1 deflation.
// Contract kept from the reference: sigma and U are indexed in place (the caller pre-sizes sigma to >= dim and U to
// rows x >= dim, image_compression/src/rSVD.cpp:111-113); V is ASSIGNED as cols x dim with the right singular vectors in
// COLUMNS (V = VT.transpose(), :27,54); A is deflated in place (A -= sigma u v^T, :21,48).
// Difference: the reference divides by sigma even when it is 0 (u = A v / sigma); here the iteration stops at
// sigma < 1e-12 like SVD<Power> (include/SVD_class.hpp:198) and the remaining entries stay zero.
#ifndef SVD_H
#define SVD_H

#include "PowerMethod.hpp"

inline void singularValueDecomposition(Mat_m& A, Vec_v& sigma, Mat_m& U, Mat_m& V, const int dim, uint64_t seed = 0x5eedULL) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols(), k = m < n ? m : n;
  if (dim < 0 || dim > k || sigma.size() < dim || U.rows() != m || U.cols() < dim) throw std::invalid_argument("singularValueDecomposition: bad dim / output sizes");
  Mat_m Uf(m, m), Vc(n, dim > 0 ? dim : 1); Vec_v Sf(k);
  int found = 0;
  rsvdb::check(c, rsvdb_svd_host(c, A.data(), m, n, m, RSVDB_SVD_POWER, dim, seed, Uf.data(), m, Sf.data(), Vc.data(), n, &found));
  Mat_m Vn = Mat_m::Zero(n, dim);
  for (int i = 0; i < dim; ++i) {
    sigma(i) = i < found ? Sf(i) : 0.0;
    for (std::ptrdiff_t r = 0; r < m; ++r) U(r, i) = i < found ? Uf(r, i) : 0.0;
    for (std::ptrdiff_t r = 0; r < n; ++r) Vn(r, i) = i < found ? Vc(r, i) : 0.0;
  }
  // in-place deflation of A, as a caller of the reference would observe it
  for (int i = 0; i < found && i < dim; ++i)
    for (std::ptrdiff_t cc = 0; cc < n; ++cc) { const double sv = Sf(i) * Vc(cc, i); for (std::ptrdiff_t r = 0; r < m; ++r) A(r, cc) -= Uf(r, i) * sv; }
  V = Vn;
}
inline void singularValueDecomposition_mpi(Mat_m& A, Vec_v& sigma, Mat_m& U, Mat_m& V, const int dim) { singularValueDecomposition(A, sigma, U, V, dim); }

#endif  // SVD_H
