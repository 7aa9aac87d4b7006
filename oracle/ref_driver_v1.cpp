// TEST INFRASTRUCTURE (oracle) -- not product code.  Nothing under rsvd_kamaneh_raganato_terrana_b200/ may use this.
//
// C entry points around the reference's OLDER first-party API, image_compression/src/{rSVD,SVD,PowerMethod,QR,
// matrixOperations,image_com}.cpp, compiled from where they lie under /root/reference by oracle/Makefile (target ref_v1 ->
// oracle/_ref/libref_imgcomp.so) against oracle/eigen_shim and the stb headers the reference vendors itself
// (image_compression/lib/).  A separate library because these translation units define functions with the same names as the
// newer src/*.cpp (rSVD, intermediate_step, ...).  No reference source is copied: this file includes the reference's own
// headers and forwards raw column-major buffers.
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include "rSVD.hpp"         // /root/reference/image_compression/include/rSVD.hpp  (5-argument rSVD, q = 1, Givens QR, power SVD)
#include "SVD.hpp"          // singularValueDecomposition
#include "PowerMethod.hpp"  // powerMethod
#include "QR.hpp"           // QRFullDecomposition / QRReducedDecomposition
#define private public      // the Image class keeps its matrices private and has no getters; the pin needs to read them
#include "image_comp.hpp"   // Image
#undef private

namespace {
using Mat = Eigen::MatrixXd; using Vec = Eigen::VectorXd;
Mat from_buf(const double* p, long r, long c) { Mat m(r, c); std::memcpy(m.data(), p, sizeof(double) * r * c); return m; }
void to_buf(const Mat& m, double* p) { std::memcpy(p, m.data(), sizeof(double) * m.rows() * m.cols()); }
struct Quiet {
  std::streambuf* old; std::streambuf* olde; std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())), olde(std::cerr.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); std::cerr.rdbuf(olde); }
};
}  // namespace

extern "C" {

// intermediate_step(A, Q, Omega, l, q) -- image_compression/include/rSVD.hpp:36, src/rSVD.cpp:7-37 (Givens QR inside)
void ref1_intermediate_step(const double* A, long m, long n, const double* Omega, int l, int q, double* Q) {
  Mat a = from_buf(A, m, n), om = from_buf(Omega, n, l), qq = Mat::Zero(m, l);
  intermediate_step(a, qq, om, l, q);
  to_buf(qq, Q);
}

// rSVD(A, U, S, V, l) -- include/rSVD.hpp:45, src/rSVD.cpp:77-118.  Omega is drawn inside from std::random_device (:92-101);
// it cannot be injected, so callers pin Omega-independent quantities only.  U m x l, S l, V n x l.
void ref1_rsvd(const double* A, long m, long n, int l, double* U, double* S, double* V, long* dims) {
  Quiet quiet;
  Mat a = from_buf(A, m, n), u = Mat::Zero(m, l), v = Mat::Zero(l, n); Vec s = Vec::Zero(l);
  rSVD(a, u, s, v, l);
  to_buf(u, U); to_buf(s, S); to_buf(v, V);
  dims[0] = u.rows(); dims[1] = u.cols(); dims[2] = s.size(); dims[3] = v.rows(); dims[4] = v.cols();
}

// singularValueDecomposition(A, sigma, U, V, dim) -- include/SVD.hpp:24, src/SVD.cpp:30-55.  sigma must be pre-sized (indexed in
// place), U rows x dim pre-sized, V is assigned cols x dim.  A is deflated in place (a copy here).
void ref1_svd(const double* A, long m, long n, int dim, double* S, double* U, double* V) {
  Quiet quiet;
  Mat a = from_buf(A, m, n), u = Mat::Zero(m, dim), v; Vec s = Vec::Zero(dim);
  singularValueDecomposition(a, s, u, v, dim);
  to_buf(s, S); to_buf(u, U); to_buf(v, V);
}

// powerMethod(A, B, sigma, u, v) -- include/PowerMethod.hpp:31, src/PowerMethod.cpp:3-43; B = A^T A formed like src/SVD.cpp:40.
void ref1_power_method(const double* A, long m, long n, double* sigma, double* u, double* v) {
  Mat a = from_buf(A, m, n); Mat b = a.transpose() * a; Vec uu = Vec::Zero(m), vv = Vec::Zero(n);
  powerMethod(a, b, *sigma, uu, vv); to_buf(uu, u); to_buf(vv, v);
}

// QRFullDecomposition<double>(A).decompose(Q, R) / QRReducedDecomposition<double> -- include/QR.hpp:32-52, src/QR.cpp:45-99.
void ref1_qr(const double* A, long m, long n, int reduced, double* Q, double* R) {
  Mat a = from_buf(A, m, n), q, r;
  if (reduced) { QRReducedDecomposition<double> d(a); d.decompose(q, r); }
  else { QRFullDecomposition<double> d(a); d.decompose(q, r); }
  to_buf(q, Q); to_buf(r, R);
}

// The Image flow of image_compression/main/main.cpp:44-69 on one rank: load(file) [stb, 1 grey channel] -> downscale(scale) ->
// normalize() -> compress(k) [rSVD with l = k + 10, :288-317] -> reconstruct() -> get_compression_ratio().
// out_norm: the normalised image matrix (rows x cols as the class holds it: width x height, image_com.cpp:40), lo / hi: the
// original range, S: l singular values, recon: U diag(S) V^T.  dims = {rows, cols, l}.  Returns -1 if the file did not load.
int ref1_image_flow(const char* path, int scale, int k, double* out_norm, double* lo, double* hi, double* S, double* recon,
                    double* ratio, long* dims) {
  Quiet quiet;
  Image img;
  img.originalWidth = 0; img.originalHeight = 0;
  img.load(path);
  if (img.image_matrix.size() == 0) return -1;
  if (scale > 1) img.downscale(scale);
  img.normalize();
  to_buf(img.image_matrix, out_norm);
  *lo = img.original_min; *hi = img.original_max;
  img.compress(k);
  to_buf(img.singular, S);
  to_buf(img.reconstruct(), recon);
  *ratio = img.get_compression_ratio();
  dims[0] = img.image_matrix.rows(); dims[1] = img.image_matrix.cols(); dims[2] = img.degree;
  return 0;
}

}  // extern "C"
