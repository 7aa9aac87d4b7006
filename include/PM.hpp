// Drop-in for reference include/PM.hpp: PM(A, B, sigma, u, v) (src/PM.cpp:4-81) -- dominant singular triplet by power
// iteration.  B = A^T A is accepted for signature compatibility and ignored: the GPU iteration is x <- A^T (A x).
#ifndef PM_H
#define PM_H

#include "rsvdb_dense.hpp"

inline void PM(Mat_m& A, Mat_m& /*B*/, double& sigma, Vec_v& u, Vec_v& v, uint64_t seed = 0x5eedULL) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols();
  Vec_v un(m), vn(n);
  rsvdb::check(c, rsvdb_pm_host(c, A.data(), m, n, m, seed, &sigma, un.data(), vn.data()));
  u = un; v = vn;
}

#endif
