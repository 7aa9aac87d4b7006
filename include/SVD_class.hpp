// Drop-in for reference include/SVD_class.hpp: enum class SVDMethod and template<SVDMethod> class SVD, same members
// and output shapes, computed by librsvdb.so (rsvdb_svd_host) instead of the host loops.
#ifndef SVD_CLASS_HPP
#define SVD_CLASS_HPP

// the standard headers the reference's SVD_class.hpp pulls in for its callers (include/SVD_class.hpp:4-21), so that
// callers relying on those transitive includes (PCA/tests/pca_test.cpp uses std::ifstream and MPI_Init through it) compile unchanged
#include <algorithm>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iostream>
#include <limits>
#include <queue>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>
#if defined(__has_include)
#if __has_include(<mpi.h>)
#include <mpi.h>
#endif
#endif

#include "rsvdb_dense.hpp"
#include "JacobiOperations.hpp"
#include "Jacobi_Class.hpp"
#include "PM.hpp"

// reference include/SVD_class.hpp:28-32
enum class SVDMethod { Jacobi, Power, ParallelJacobi };

// reference include/SVD_class.hpp:35-71
template <SVDMethod method>
class SVD {
 public:
  SVD(const Mat_m& data, const int& r = 0) : data_(data), r_(r) {}   // copies its input like the reference (:74-75)

  // :79-97.  (The reference also prints "Entering compute() method" etc. to stdout; the drop-in stays quiet.)
  void compute() {
    rsvdb_ctx* c = rsvdb::default_context();
    const std::ptrdiff_t m = data_.rows(), n = data_.cols(), k = m < n ? m : n;
    int found = 0;
    if (method == SVDMethod::Power) {
      // U_ m x m identity-completed, S_ min(m,n), V_ n x n with the right singular vectors in ROWS (:83,213-214)
      const std::ptrdiff_t dim = r_ ? r_ : k;
      Mat_m U(m, m), Vc(n, dim); Vec_v S(k);
      rsvdb::check(c, rsvdb_svd_host(c, data_.data(), m, n, m, static_cast<int>(method), r_, seed_, U.data(), m, S.data(), Vc.data(), n, &found));
      Mat_m V = Mat_m::Identity(n, n);
      for (std::ptrdiff_t i = 0; i < found && i < dim; ++i) for (std::ptrdiff_t j = 0; j < n; ++j) V(i, j) = Vc(j, i);
      if (found < dim) {                         // conservativeResize on the sigma < 1e-12 early exit (:198-209)
        const std::ptrdiff_t f = found ? found : 1;
        Mat_m U2(m, f), V2(n, f); Vec_v S2(f);
        for (std::ptrdiff_t j = 0; j < found; ++j) { S2(j) = S(j); for (std::ptrdiff_t i = 0; i < m; ++i) U2(i, j) = U(i, j); for (std::ptrdiff_t i = 0; i < n; ++i) V2(i, j) = V(i, j); }
        U_ = U2; S_ = S2; V_ = V2;
      } else { U_ = U; S_ = S; V_ = V; }
    } else {
      // U_ m x k, S_ k, V_ n x k (:105-107,158-178)
      Mat_m U(m, k), V(n, k); Vec_v S(k);
      rsvdb::check(c, rsvdb_svd_host(c, data_.data(), m, n, m, static_cast<int>(method), r_, seed_, U.data(), m, S.data(), V.data(), n, &found));
      U_ = U; S_ = S; V_ = V;
    }
  }

  Mat_m getU() const { return U_; }
  Vec_v getS() const { return S_; }
  Mat_m getV() const { return V_; }
  void setSeed(uint64_t s) { seed_ = s; }     // additive: the Power back-end's start vector (std::random_device in the reference)

 private:
  Mat_m U_; Vec_v S_; Mat_m V_; Mat_m data_; int r_; uint64_t seed_ = 0;

 protected:
  void setData(const Mat_m& data) { data_ = data; }   // :67-70 (used by PCA<method> : SVD<method>)
  // additive, for PCA<method>: the fused centre + SVD entry point (rsvdb_pca_host) delivers the factors directly
  void setResults(const Mat_m& U, const Vec_v& S, const Mat_m& V) { U_ = U; S_ = S; V_ = V; }
};

#endif  // SVD_CLASS_HPP
