// Drop-in for reference include/Jacobi_Class.hpp / src/Jacobi_Class.cpp: the JacobiRotation value type.  A handful of
// scalar operations per call: it stays a host-side inline type; the rotations of the SVD itself
// run inside the GPU Jacobi kernel.
#ifndef JACOBIROTATION_H
#define JACOBIROTATION_H

#include <cmath>
#include <iostream>
#include <limits>

#include "rsvdb_dense.hpp"

class JacobiRotation {
 public:
  JacobiRotation() : c_(1), s_(0) { set_identity(); }                                   // src/Jacobi_Class.cpp:5
  JacobiRotation(double c, double s) : c_(c), s_(s) { mat_[0] = c; mat_[1] = -s; mat_[2] = s; mat_[3] = c; }   // :7-10  [c s; -s c]
  double getC() const { return c_; }
  double getS() const { return s_; }
  void setC(double c) { c_ = c; }
  void setS(double s) { s_ = s; }
  // 2 x 2 [c s; -s c] (:28-33)
  Mat_m getMatrix() const { Mat_m m(2, 2); m(0, 0) = c_; m(0, 1) = s_; m(1, 0) = -s_; m(1, 1) = c_; return m; }
  // getMatrix() * vec (:35-37)
  Vec_v apply(const Vec_v& vec) const { Vec_v r(2); r(0) = c_ * vec(0) + s_ * vec(1); r(1) = -s_ * vec(0) + c_ * vec(1); return r; }
  // :39-60
  bool makeJacobi(double x, double y, double z) {
    const double deno = 2 * std::abs(y);
    if (deno < (std::numeric_limits<double>::min)()) { setC(1.0); setS(0.0); return false; }
    const double tau = (x - z) / deno, w = std::sqrt(tau * tau + 1);
    const double t = tau > 0 ? 1 / (tau + w) : 1 / (tau - w);
    const double sgn = t > 0 ? 1 : -1, n = 1 / std::sqrt(t * t + 1);
    setS(-sgn * (y / std::abs(y)) * std::abs(t) * n); setC(n);
    return true;
  }
  // like the reference (include/Jacobi_Class.hpp:26-30) this transposes the stored 2 x 2 only, not c / s
  JacobiRotation transpose() { JacobiRotation t; t.mat_[0] = mat_[0]; t.mat_[1] = mat_[2]; t.mat_[2] = mat_[1]; t.mat_[3] = mat_[3]; return t; }
  Mat_m operator*(const JacobiRotation& o) const {
    Mat_m r(2, 2);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) r(i, j) = mat_[i] * o.mat_[2 * j] + mat_[2 + i] * o.mat_[2 * j + 1];
    return r;
  }
  void setMat(const Mat_m& m) { mat_[0] = m(0, 0); mat_[1] = m(1, 0); mat_[2] = m(0, 1); mat_[3] = m(1, 1); }
  void printMatrix() const { std::cout << c_ << " " << s_ << "\n" << -s_ << " " << c_ << std::endl; }
 private:
  void set_identity() { mat_[0] = 1; mat_[1] = 0; mat_[2] = 0; mat_[3] = 1; }
  double c_, s_;
  double mat_[4];   // column-major 2 x 2
};

#endif  // JACOBIROTATION_H
