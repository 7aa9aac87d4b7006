// TEST INFRASTRUCTURE (oracle) -- not product code.  Nothing under rsvd_kamaneh_raganato_terrana_b200/ may use this.
//
// C entry points around the reference's OWN first-party translation units, which oracle/Makefile compiles from
// where they lie under /root/reference (src/rSVD.cpp, src/JacobiOperations.cpp, src/Jacobi_Class.cpp, src/PM.cpp,
// src/QR.cpp, src/matrixOperations.cpp + include/*.hpp) against oracle/eigen_shim (Eigen and MPI are absent from
// this image; see eigen_shim/Eigen/Dense for what is restated).  No reference source is copied here: this file only
// declares the reference's public functions through the reference's own headers and forwards raw column-major
// buffers to them.
#include <cstring>
#include <sstream>
#include <iostream>
#include "rSVD.hpp"        // /root/reference/include/rSVD.hpp
#include "SVD_class.hpp"   // /root/reference/include/SVD_class.hpp
#include "QR.hpp"          // /root/reference/include/QR.hpp
#include "PM.hpp"          // /root/reference/include/PM.hpp
#include <mpi.h>           // oracle/eigen_shim/mpi.h
#include "PCA_class.hpp"   // /root/reference/PCA/include/PCA_class.hpp
#include "POD.hpp"         // /root/reference/POD/ParametricDiffusion1D/src/POD.hpp

namespace {
Mat_m from_buf(const double* p, long r, long c) { Mat_m m(r, c); std::memcpy(m.data(), p, sizeof(double) * r * c); return m; }
void to_buf(const Eigen::MatrixXd& m, double* p) { std::memcpy(p, m.data(), sizeof(double) * m.rows() * m.cols()); }
struct Quiet {  // the reference prints from library code (include/SVD_class.hpp:80,87); keep test logs readable
  std::streambuf* old; std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

// PCA<method>(data, normalize) -- PCA/include/PCA_class.hpp:12-207.  method: 0 Jacobi, 2 ParallelJacobi.  k = min(m, n).
// Outputs: expl_var[k], ratio[k], scores[m*k], loadings[n*k], mean[n] (= reconstructFromPCA(0), mean_ has no getter),
// proj[pr*k] = projectToPCA(P), recon[pr*n] = reconstructFromPCA(proj), orth = checkOrthogonality().
// Returns -1 on the reference's std::invalid_argument (assertDataValid).
template <SVDMethod M>
static int ref_pca_t(const Mat_m& d, bool normalize, const Mat_m& P, double* ev, double* ratio, double* scores, double* loadings,
                     double* mean, double* proj, double* recon, double* orth) {
  try {
    PCA<M> pca(d, normalize);
    to_buf(pca.explainedVariance(), ev); to_buf(pca.explainedVarianceRatio(), ratio);
    to_buf(pca.scores(), scores); to_buf(pca.loadings(), loadings);
    const long k = pca.loadings().cols();
    to_buf(pca.reconstructFromPCA(Mat_m::Zero(1, k)), mean);
    Mat_m pr = pca.projectToPCA(P); to_buf(pr, proj); to_buf(pca.reconstructFromPCA(pr), recon);
    *orth = pca.checkOrthogonality();
  } catch (const std::invalid_argument&) { return -1; }
  return 0;
}

extern "C" {

// intermediate_step(A, Q, Omega, l, q)  -- include/rSVD.hpp:13, src/rSVD.cpp:57-70.  Q is m x l.
void ref_intermediate_step(const double* A, long m, long n, const double* Omega, int l, int q, double* Q) {
  Mat_m a = from_buf(A, m, n), om = from_buf(Omega, n, l), qq = Mat_m::Zero(m, l);
  intermediate_step(a, qq, om, l, q);
  to_buf(qq, Q);
}

// rSVD(A, U, S, V, l, method) -- include/rSVD.hpp:14, src/rSVD.cpp:72-133, with Omega delivered through the MPI stub's
// one-shot Bcast override (see eigen_shim/mpi.h).  method: 0 Jacobi, 1 Power, 2 ParallelJacobi (include/SVD_class.hpp:28-32).
// Output shapes are whatever the reference assigns; they are reported through dims = {Ur,Uc,Sn,Vr,Vc}.
// Caller provides U (m*max(l,n)), S (max(l,n)), V (n*n) capacity.
int ref_rsvd(const double* A, long m, long n, const double* Omega, int l, int method, double* U, double* S, double* V, long* dims) {
  Quiet quiet;
  Mat_m a = from_buf(A, m, n), u, v; Vec_v s;
  Mat_m om = from_buf(Omega, n, l);
  oracle_mpi_set_bcast_override(om.data(), static_cast<std::size_t>(n) * l);
  try {
    rSVD(a, u, s, v, l, static_cast<SVDMethod>(method));
  } catch (const std::invalid_argument&) {
    return -1;
  }
  to_buf(u, U); to_buf(s, S); to_buf(v, V);
  dims[0] = u.rows(); dims[1] = u.cols(); dims[2] = s.size(); dims[3] = v.rows(); dims[4] = v.cols();
  return 0;
}

// SVD<method>(data, r).compute(); getU/getS/getV -- include/SVD_class.hpp:35-97.
int ref_svd(const double* A, long m, long n, int method, int r, double* U, double* S, double* V, long* dims) {
  Quiet quiet;
  Mat_m a = from_buf(A, m, n), u, v; Vec_v s;
  switch (method) {
    case 0: { SVD<SVDMethod::Jacobi> svd(a, r); svd.compute(); u = svd.getU(); s = svd.getS(); v = svd.getV(); break; }
    case 1: { SVD<SVDMethod::Power> svd(a, r); svd.compute(); u = svd.getU(); s = svd.getS(); v = svd.getV(); break; }
    case 2: { SVD<SVDMethod::ParallelJacobi> svd(a, r); svd.compute(); u = svd.getU(); s = svd.getS(); v = svd.getV(); break; }
    default: return -1;
  }
  to_buf(u, U); to_buf(s, S); to_buf(v, V);
  dims[0] = u.rows(); dims[1] = u.cols(); dims[2] = s.size(); dims[3] = v.rows(); dims[4] = v.cols();
  return 0;
}

// qr_decomposition_reduced / qr_decomposition_full -- include/QR.hpp:15-16, src/QR.cpp:22-80.
void ref_qr_reduced(const double* A, long m, long n, double* Q, double* R) {
  Mat_m a = from_buf(A, m, n), q, r; qr_decomposition_reduced(a, q, r); to_buf(q, Q); to_buf(r, R);
}
void ref_qr_full(const double* A, long m, long n, double* Q, double* R) {
  Mat_m a = from_buf(A, m, n), q, r; qr_decomposition_full(a, q, r); to_buf(q, Q); to_buf(r, R);
}

// PM(A, B, sigma, u, v) -- include/PM.hpp:18, src/PM.cpp:4-81.  B = A^T A is formed by the caller in the reference
// (include/SVD_class.hpp:193); here it is formed with the shim product.
void ref_pm(const double* A, long m, long n, double* sigma, double* u, double* v) {
  Mat_m a = from_buf(A, m, n); Mat_m b = a.transpose() * a; Vec_v uu = Vec_v::Zero(m), vv = Vec_v::Zero(n);
  PM(a, b, *sigma, uu, vv); to_buf(uu, u); to_buf(vv, v);
}

// manualMatrixMultiply -- include/matrixOperations.hpp:14, src/matrixOperations.cpp:7-28.  Returns -1 on the
// reference's std::invalid_argument.
int ref_manual_matmul(const double* A, long m, long k, const double* B, long k2, long n, double* C) {
  try { Mat c = manualMatrixMultiply(from_buf(A, m, k), from_buf(B, k2, n)); to_buf(c, C); return 0; }
  catch (const std::invalid_argument&) { return -1; }
}

// makeJacobi / real_2x2_jacobi_svd -- src/Jacobi_Class.cpp:39-60, src/JacobiOperations.cpp:25-88.
int ref_make_jacobi(double x, double y, double z, double* c, double* s) {
  JacobiRotation r; bool ok = r.makeJacobi(x, y, z); *c = r.getC(); *s = r.getS(); return ok ? 1 : 0;
}
void ref_real_2x2_jacobi_svd(const double* M4_colmajor, double* cl, double* sl, double* cr, double* sr) {
  Mat_m m = from_buf(M4_colmajor, 2, 2); real_2x2_jacobi_svd(m, *cl, *sl, *cr, *sr, 0, 1);
}


// PCA<method> through the reference's own class (see ref_pca_t above)
int ref_pca(const double* data, long m, long n, int normalize, int method, const double* P, long pr, double* ev, double* ratio,
            double* scores, double* loadings, double* mean, double* proj, double* recon, double* orth) {
  Quiet quiet;
  Mat_m d = from_buf(data, m, n), p = from_buf(P, pr, n);
  if (method == 0) return ref_pca_t<SVDMethod::Jacobi>(d, normalize != 0, p, ev, ratio, scores, loadings, mean, proj, recon, orth);
  if (method == 2) return ref_pca_t<SVDMethod::ParallelJacobi>(d, normalize != 0, p, ev, ratio, scores, loadings, mean, proj, recon, orth);
  return -2;
}


// POD(S, r, svd_type) / POD(S, r, tol, svd_type) / POD(S, Xh, r, tol, svd_type) / POD(S, Xh, D, r, tol, svd_type)
// -- POD/ParametricDiffusion1D/src/POD.hpp:26-38, POD.cpp:11-40.  variant 0..3 picks the constructor.  Omega (optional):
// delivered to rSVD's generateOmega through the MPI stub's one-shot Bcast override (svd_type 3-5).
// W / sigma capacities are the caller's; dims = {W rows, W cols, sigma size}.
int ref_pod(int variant, const double* S, long Nh, long ns, const double* Xh, const double* D, int r, double tol, int svd_type,
            const double* Omega, long om_rows, double* W, double* sigma, long* dims) {
  Quiet quiet;
  Mat_m s = from_buf(S, Nh, ns), xh, d, om;
  if (variant >= 2) xh = from_buf(Xh, Nh, Nh);
  if (variant == 3) d = from_buf(D, ns, ns);
  if (Omega) { om = from_buf(Omega, om_rows, r); oracle_mpi_set_bcast_override(om.data(), static_cast<std::size_t>(om_rows) * r); }
  POD* p = nullptr;
  switch (variant) {
    case 0: p = new POD(s, r, svd_type); break;
    case 1: p = new POD(s, r, tol, svd_type); break;
    case 2: p = new POD(s, xh, r, tol, svd_type); break;
    case 3: p = new POD(s, xh, d, r, tol, svd_type); break;
    default: return -1;
  }
  to_buf(p->W, W); to_buf(p->sigma, sigma);
  dims[0] = p->W.rows(); dims[1] = p->W.cols(); dims[2] = p->sigma.size();
  delete p;
  return 0;
}

}  // extern "C"
