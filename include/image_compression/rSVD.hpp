// Drop-in for reference image_compression/include/rSVD.hpp (image_compression/src/rSVD.cpp:7-186): the OLDER rSVD API --
// rSVD(A, U, S, V, l) with q = 1 power iteration (:103), Givens QR, power-method small SVD; V comes back n x min(l, n)
// with the right singular vectors in COLUMNS.  Include this INSTEAD of ../rSVD.hpp (both define intermediate_step).
// SURVEY.md 8(f) rank 1.
#ifndef rSVD_V1_H
#define rSVD_V1_H

#include "../rsvdb_dense.hpp"
#include "SVD.hpp"

// image_compression/src/rSVD.cpp:7-37 (same sequence as src/rSVD.cpp:57-70; l and q by non-const reference there)
inline void intermediate_step(Mat_m& A, Mat_m& Q, Mat_m& Omega, int& l, int& q) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols();
  Mat_m Qn(m, l);
  rsvdb::check(c, rsvdb_intermediate_step_host(c, A.data(), m, n, m, Omega.data(), Omega.rows(), l, q, Qn.data(), m));
  Q = Qn;
}
// :39-74.  (The reference's MPI variant factorises A instead of Y0, a bug -- SURVEY.md 3.5; this is the serial semantics.)
inline void intermediate_step_mpi(Mat_m& A, Mat_m& Q, Mat_m& Omega, int& l, int& q) { intermediate_step(A, Q, Omega, l, q); }

// :77-118: Omega ~ N(0,1) (std::random_device there, seeded here), q = 1, singularValueDecomposition(B, S, Utilde, V, min_dim)
inline void rSVD(Mat_m& A, Mat_m& U, Vec_v& S, Mat_m& V, int l, uint64_t seed = 0x5eedULL) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols(), k = l < n ? l : n;
  Mat_m Un(m, k), Vn(n, k); Vec_v Sn(k);
  rsvdb::check(c, rsvdb_rsvd_host(c, A.data(), m, n, m, nullptr, 0, seed, l, /*q=*/1, RSVDB_SVD_POWER, Un.data(), m, Sn.data(), Vn.data(), n));
  U = Un; V = Vn;
  if (S.size() < k) S = Sn; else for (std::ptrdiff_t i = 0; i < k; ++i) S(i) = Sn(i);   // the reference indexes S in place
}
inline void rSVD_mpi(Mat_m& A, Mat_m& U, Vec_v& S, Mat_m& V, int l) { rSVD(A, U, S, V, l); }

#endif
