// Matrix types for the C++ drop-in headers.
//
// The reference types its API with Eigen (`using Mat_m = Eigen::MatrixXd; using Vec_v = Eigen::VectorXd;`,
// reference include/rSVD.hpp:9-10).  When Eigen is available these headers use it, so existing callers compile
// unchanged.  Where it is not (this image), a minimal column-major container with the same storage layout and the
// accessors the reference's callers use (rows/cols/size/data/operator()/Zero/Identity/Constant/
// resize) stands in.  It does NO arithmetic: every product, factorisation and rotation happens in librsvdb.so on the GPU.
#ifndef RSVDB_DENSE_HPP
#define RSVDB_DENSE_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "rsvdb.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(RSVDB_NO_EIGEN)
#define RSVDB_HAVE_EIGEN 1
#endif
#endif

#ifdef RSVDB_HAVE_EIGEN
#include <Eigen/Dense>
using Mat_m = Eigen::MatrixXd;
using Vec_v = Eigen::VectorXd;
#else
namespace rsvdb {
using Index = std::ptrdiff_t;
class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(Index r, Index c) : r_(r), c_(c), d_(static_cast<size_t>(r * c), 0.0) {}
  static Matrix Constant(Index r, Index c, double v) { Matrix m(r, c); for (auto& x : m.d_) x = v; return m; }
  static Matrix Zero(Index r, Index c) { return Matrix(r, c); }
  static Matrix Identity(Index r, Index c) { Matrix m(r, c); for (Index i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0; return m; }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double& operator()(Index i, Index j) { return d_[static_cast<size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return d_[static_cast<size_t>(i + j * r_)]; }
  double& operator()(Index i) { return d_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return d_[static_cast<size_t>(i)]; }
  double& operator[](Index i) { return d_[static_cast<size_t>(i)]; }
  double operator[](Index i) const { return d_[static_cast<size_t>(i)]; }
  void resize(Index r, Index c) { r_ = r; c_ = c; d_.assign(static_cast<size_t>(r * c), 0.0); }
 protected:
  Index r_, c_;
  std::vector<double> d_;
};
class Vector : public Matrix {
 public:
  Vector() : Matrix(0, 1) {}
  explicit Vector(Index n) : Matrix(n, 1) {}
  static Vector Zero(Index n) { return Vector(n); }
  void resize(Index n) { Matrix::resize(n, 1); }
};
}  // namespace rsvdb
using Mat_m = rsvdb::Matrix;
using Vec_v = rsvdb::Vector;
#endif

namespace rsvdb {

// One engine context per process (GPU 0 unless RSVDB_DEVICE is set).  The reference needs MPI_Init before any call
// (generateOmega / PM query MPI_COMM_WORLD); here nothing needs initialising.
inline rsvdb_ctx* default_context() {
  static rsvdb_ctx* ctx = [] {
    rsvdb_ctx* c = nullptr;
    int dev = 0;
    if (const char* e = std::getenv("RSVDB_DEVICE")) dev = std::atoi(e);
    const int rc = rsvdb_create(&c, dev);
    if (rc != RSVDB_OK) throw std::runtime_error("rsvdb_create failed (" + std::to_string(rc) + "): no B200 / sm_100 GPU visible; there is no CPU fallback");
    return c;
  }();
  return ctx;
}

// The reference signals errors with std::invalid_argument (src/rSVD.cpp:122-123, src/matrixOperations.cpp:8-11);
// the C ABI returns codes, the wrappers re-throw.
inline void check(rsvdb_ctx* c, int rc) {
  if (rc == RSVDB_OK) return;
  const std::string msg = rsvdb_last_error(c);
  if (rc == RSVDB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error("rsvdb error " + std::to_string(rc) + ": " + msg);
}

inline void resize_matrix(Mat_m& m, std::ptrdiff_t r, std::ptrdiff_t c) { m.resize(r, c); }
inline void resize_vector(Vec_v& v, std::ptrdiff_t n) { v.resize(n); }

}  // namespace rsvdb
#endif
