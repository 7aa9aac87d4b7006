// Drop-in for reference include/JacobiOperations.hpp / src/JacobiOperations.cpp:6-103,120-203: the plane-rotation
// appliers and the real 2 x 2 SVD.  In the reference these are the inner loop of the Jacobi SVD; in this engine that
// loop runs inside one GPU kernel (csrc/jacobi.cu) and never calls back to the host.  The functions remain available
// as host-side scalar utilities with the reference's exact signatures and arithmetic, for
// callers that use them directly on small matrices.
#ifndef JACOBISVD_H
#define JACOBISVD_H

#include <cmath>
#include <limits>

#include "rsvdb_dense.hpp"
#include "Jacobi_Class.hpp"

inline void applyOnTheLeft(Mat_m& matrix, int p, int q, double c, double s) {          // :6-14
  for (std::ptrdiff_t i = 0; i < matrix.cols(); ++i) {
    const double xi = matrix(p, i), yi = matrix(q, i);
    matrix(p, i) = c * xi + s * yi; matrix(q, i) = -s * xi + c * yi;
  }
}
inline void applyOnTheRight(Mat_m& matrix, int p, int q, double c, double s) {         // :16-24
  for (std::ptrdiff_t i = 0; i < matrix.rows(); ++i) {
    const double xi = matrix(i, p), yi = matrix(i, q);
    matrix(i, p) = c * xi + (-s) * yi; matrix(i, q) = s * xi + c * yi;
  }
}
namespace rsvdb {
inline void real_2x2_svd(double m00, double m01, double m10, double m11, double deno_floor, double& cl, double& sl, double& cr, double& sr) {
  const double t = m00 + m11, d = m10 - m01;
  double rc, rs;
  if (d == 0) { rs = 0.0; rc = 1.0; } else { const double u = t / d, tmp = std::sqrt(1.0 + u * u); rs = 1 / tmp; rc = u / tmp; }
  const double a00 = rc * m00 + rs * m10, a01 = rc * m01 + rs * m11, a11 = -rs * m01 + rc * m11;
  const double deno = 2 * std::abs(a01);
  if (deno < deno_floor) { cr = 1; sr = 0; }
  else {
    const double tau = (a00 - a11) / deno, w = std::sqrt(tau * tau + 1);
    const double t2 = tau > 0 ? 1 / (tau + w) : 1 / (tau - w);
    const double sgn = t2 > 0 ? 1 : -1, n = 1 / std::sqrt(t2 * t2 + 1);
    sr = -sgn * (a01 / std::abs(a01)) * std::abs(t2) * n; cr = n;
  }
  cl = rc * cr + rs * sr; sl = rc * (-sr) + rs * cr;
}
}  // namespace rsvdb
inline void real_2x2_jacobi_svd(Mat_m& matrix, double& c_left, double& s_left, double& c_right, double& s_right, int p, int q) {   // :25-88
  rsvdb::real_2x2_svd(matrix(p, p), matrix(p, q), matrix(q, p), matrix(q, q), (std::numeric_limits<double>::min)(), c_left, s_left, c_right, s_right);
}
inline bool svd_precondition_2x2_block_to_be_real(Mat_m& m_workMatrix, int p, int q, double maxDiagEntry) {                      // :89-103
  const double eps = std::numeric_limits<double>::epsilon();
  return !(std::abs(m_workMatrix(p, q)) < maxDiagEntry * eps && std::abs(m_workMatrix(q, p)) < maxDiagEntry * eps);
}
// "_par" twins (:120-203): same arithmetic (OpenMP only split the loops); the 2 x 2 SVD uses deno < 1e-10 (:168)
inline void applyOnTheLeft_par(Mat_m& matrix, size_t p, size_t q, double c, double s) { applyOnTheLeft(matrix, static_cast<int>(p), static_cast<int>(q), c, s); }
inline void applyOnTheRight_par(Mat_m& matrix, size_t p, size_t q, double c, double s) { applyOnTheRight(matrix, static_cast<int>(p), static_cast<int>(q), c, s); }
inline void real_2x2_jacobi_svd_par(Mat_m& matrix, double& c_left, double& s_left, double& c_right, double& s_right, size_t p, size_t q) {
  rsvdb::real_2x2_svd(matrix(p, p), matrix(p, q), matrix(q, p), matrix(q, q), 1e-10, c_left, s_left, c_right, s_right);
}

#endif  // JACOBISVD_H
