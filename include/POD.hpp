// Drop-in for reference POD/ParametricDiffusion1D/src/POD.hpp: class POD with the same four constructors and the public
// members W (POD modes) and sigma.  Each constructor is ONE call into librsvdb.so (rsvdb_pod_host): the snapshot matrix
// is uploaded once; the correlation matrix, the SVD / rSVD the svd_type selects (perform_SVD, POD.cpp:42-114), the mode
// recovery and the energy criterion run on the device.  The progress prints of the reference are dropped.
#ifndef POD_H
#define POD_H

#include <cstdlib>
#include <iostream>
#include <tuple>

#include "SVD_class.hpp"
#include "rSVD.hpp"

class POD {
 public:
  POD() {}                                                                            // POD.cpp:4-8
  POD(Mat_m& S, const int r, const int svd_type) { run(0, S, nullptr, nullptr, r, 0.0, svd_type); }                      // naive    :11-16
  POD(Mat_m& S, const int r, const double tol, const int svd_type) { run(1, S, nullptr, nullptr, r, tol, svd_type); }   // standard :19-24
  POD(Mat_m& S, Mat_m& Xh, const int r, const double tol, const int svd_type) { run(2, S, &Xh, nullptr, r, tol, svd_type); }   // energy :27-32
  POD(Mat_m& S, Mat_m& Xh, Mat_m& D, const int r, const double tol, const int svd_type) { run(3, S, &Xh, &D, r, tol, svd_type); }   // weight :35-40

  Mat_m W;       // POD modes
  Vec_v sigma;   // singular values as the reference stores them (of the correlation matrix for the standard / energy / weight variants)

  // additive: the sketch of svd_type 3-5 is drawn on the device from this seed (std::random_device in the reference)
  static uint64_t& seed() { static uint64_t s = 0; return s; }

 private:
  void run(int variant, Mat_m& S, Mat_m* Xh, Mat_m* D, int r, double tol, int svd_type) {
    rsvdb_ctx* c = rsvdb::default_context();
    int64_t wc = 0, sl = 0;
    if (rsvdb_pod_shape(variant, S.rows(), S.cols(), r, svd_type, &wc, &sl) != 0) {
      if (svd_type < 0 || svd_type > 5) {                                             // POD.cpp:87-91
        std::cerr << "The svd_type should be in [0,5]. Check 'svd_type' in the parameter file." << std::endl;
        std::exit(EXIT_FAILURE);
      }
      throw std::invalid_argument("POD: bad argument");
    }
    Mat_m Wfull(S.rows(), wc); Vec_v sg = Vec_v::Zero(sl);
    int N = 0;
    rsvdb::check(c, rsvdb_pod_host(c, variant, S.data(), S.rows(), S.cols(), S.rows(), Xh ? Xh->data() : nullptr, Xh ? Xh->rows() : 0,
                                   D ? D->data() : nullptr, D ? D->rows() : 0, r, tol, svd_type, seed(), nullptr, 0, Wfull.data(), S.rows(),
                                   sg.data(), &N));
    Mat_m Wn(S.rows(), N);                                                            // W.conservativeResize(NoChange, N), POD.cpp:221
    for (std::ptrdiff_t j = 0; j < N; ++j) for (std::ptrdiff_t i = 0; i < S.rows(); ++i) Wn(i, j) = Wfull(i, j);
    W = Wn; sigma = sg;
  }
};

#endif  // POD_H
