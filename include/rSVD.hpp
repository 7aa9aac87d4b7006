// Drop-in for reference include/rSVD.hpp: same three functions, same argument meaning, outputs assigned (resized) like
// the reference; plus one additive overload that exposes Omega and q (hard-coded / internal in the reference).
#ifndef rSVD_H
#define rSVD_H

#include "rsvdb_dense.hpp"
#include "SVD_class.hpp"

// reference include/rSVD.hpp:13, src/rSVD.cpp:57-70
inline void intermediate_step(const Mat_m& A, Mat_m& Q, const Mat_m& Omega, int l, int q) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols();
  Mat_m Qn(m, l);
  rsvdb::check(c, rsvdb_intermediate_step_host(c, A.data(), m, n, m, Omega.data(), Omega.rows(), l, q, Qn.data(), m));
  Q = Qn;
}

// reference include/rSVD.hpp:15, src/rSVD.cpp:12-55 (N(0,1) entries; the reference seeds from std::random_device)
inline Mat_m generateOmega(int n, int l, uint64_t seed = 0x5eedULL) {
  rsvdb_ctx* c = rsvdb::default_context();
  Mat_m Om(n, l);
  rsvdb::check(c, rsvdb_generate_omega_host(c, n, l, seed, Om.data(), n));
  return Om;
}

namespace rsvdb {
inline void rsvd_impl(const Mat_m& A, Mat_m& U, Vec_v& S, Mat_m& V, int l, SVDMethod method, const double* Omega, std::ptrdiff_t ldo,
                      int q, uint64_t seed) {
  rsvdb_ctx* c = default_context();
  const int im = static_cast<int>(method);
  if (im != 0 && im != 1 && im != 2) throw std::invalid_argument("Unsupported SVD method");   // src/rSVD.cpp:122-123
  const std::ptrdiff_t m = A.rows(), n = A.cols(), k = l < n ? l : n;
  Mat_m Un(m, k), Vn(n, k); Vec_v Sn(k);
  check(c, rsvdb_rsvd_host(c, A.data(), m, n, m, Omega, ldo, seed, l, q, im, Un.data(), m, Sn.data(), Vn.data(), n));
  U = Un; S = Sn;
  if (method == SVDMethod::Power) {            // V_ is n x n with the vectors in rows (include/SVD_class.hpp:83,214)
    Mat_m Vr = Mat_m::Identity(n, n);
    for (std::ptrdiff_t i = 0; i < k; ++i) for (std::ptrdiff_t j = 0; j < n; ++j) Vr(i, j) = Vn(j, i);
    V = Vr;
  } else {
    V = Vn;
  }
}
}  // namespace rsvdb

// reference include/rSVD.hpp:14, src/rSVD.cpp:72-133 (q = 2, :83; Omega drawn internally, :81)
inline void rSVD(Mat_m& A, Mat_m& U, Vec_v& S, Mat_m& V, int l, SVDMethod method) {
  rsvdb::rsvd_impl(A, U, S, V, l, method, nullptr, 0, 2, 0x5eedULL);
}
// additive overload: caller-supplied Omega (n x l) and q
inline void rSVD(Mat_m& A, Mat_m& U, Vec_v& S, Mat_m& V, int l, SVDMethod method, const Mat_m& Omega, int q) {
  if (Omega.rows() != A.cols() || Omega.cols() != l) throw std::invalid_argument("Omega must be n x l");
  rsvdb::rsvd_impl(A, U, S, V, l, method, Omega.data(), Omega.rows(), q, 0);
}

#endif
