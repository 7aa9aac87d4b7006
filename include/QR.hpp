// Drop-in for reference include/QR.hpp (free functions, src/QR.cpp:22-80) and image_compression/include/QR.hpp
// (the QR class hierarchy, image_compression/src/QR.cpp:28-154).  The reference rotates adjacent rows with Givens
// rotations on the host (O(m^2 n) work, an explicit m x m Q); here the factorisation is Householder TSQR on the GPU
// (rsvdb_qr_host), sign-normalised to the Givens convention diag(R) >= 0.
#ifndef QR_H
#define QR_H

#include "rsvdb_dense.hpp"
#include "matrixOperations.hpp"

// src/QR.cpp:22-41: Q m x m, R m x n
inline void qr_decomposition_full(const Mat_m& A, Mat_m& Q, Mat_m& R) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols();
  Mat_m Qn(m, m), Rn(m, n);
  rsvdb::check(c, rsvdb_qr_host(c, A.data(), m, n, m, 1, Qn.data(), m, Rn.data(), m));
  Q = Qn; R = Rn;
}
// src/QR.cpp:43-80: Q m x n, R n x n (requires m >= n, :78-79)
inline void qr_decomposition_reduced(const Mat_m& A, Mat_m& Q, Mat_m& R) {
  rsvdb_ctx* c = rsvdb::default_context();
  const std::ptrdiff_t m = A.rows(), n = A.cols();
  Mat_m Qn(m, n), Rn(n, n);
  rsvdb::check(c, rsvdb_qr_host(c, A.data(), m, n, m, 0, Qn.data(), m, Rn.data(), n));
  Q = Qn; R = Rn;
}

// image_compression/include/QR.hpp:13-64.  The reference templates on Scalar and instantiates double only
// (image_compression/src/QR.cpp:157-160); A is held BY REFERENCE (:27) and must outlive the object.
template <typename Scalar>
class QRDecomposition {
  static_assert(sizeof(Scalar) == sizeof(double), "the reference instantiates double only");
 public:
  explicit QRDecomposition(const Mat_m& A) : A_(A) {}
  virtual ~QRDecomposition() {}
  virtual void decompose(Mat_m& Q, Mat_m& R) const = 0;
 protected:
  const Mat_m& getA() const { return A_; }
 private:
  const Mat_m& A_;
};
template <typename Scalar>
class QRFullDecomposition : public QRDecomposition<Scalar> {
 public:
  explicit QRFullDecomposition(const Mat_m& A) : QRDecomposition<Scalar>(A) {}
  void decompose(Mat_m& Q, Mat_m& R) const override { qr_decomposition_full(this->getA(), Q, R); }
};
template <typename Scalar>
class QRReducedDecomposition : public QRDecomposition<Scalar> {
 public:
  explicit QRReducedDecomposition(const Mat_m& A) : QRDecomposition<Scalar>(A) {}
  void decompose(Mat_m& Q, Mat_m& R) const override { qr_decomposition_reduced(this->getA(), Q, R); }
};
// The reference's MPI variant broadcasts with mismatched roots and is only meaningful at one rank;
// it returns the reduced factorisation, as does this.
template <typename Scalar>
class QRMPIDecomposition : public QRDecomposition<Scalar> {
 public:
  explicit QRMPIDecomposition(const Mat_m& A) : QRDecomposition<Scalar>(A) {}
  void decompose(Mat_m& Q, Mat_m& R) const override { qr_decomposition_reduced(this->getA(), Q, R); }
};

#endif
