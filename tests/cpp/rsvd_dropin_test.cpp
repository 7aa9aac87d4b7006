// C++ caller written against the reference's own API (the shape of reference tests/rSVD_test.cpp:54-97 and
// tests/rSVD_test2.cpp:96-124), compiled against the drop-in headers in include/ and linked to librsvdb.so.
// It reads a column-major binary matrix, runs rSVD / SVD<> / QR / PM / manualMatrixMultiply, and writes the results as
// raw doubles for the Python side of the test to compare with the oracle.
//   usage: rsvd_dropin_test <in.bin> <m> <n> <l> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "rSVD.hpp"
#include "QR.hpp"
#include "PM.hpp"
#include "matrixOperations.hpp"

static void dump(const std::string& path, const double* p, size_t n) {
  std::ofstream f(path, std::ios::binary); f.write(reinterpret_cast<const char*>(p), sizeof(double) * n);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage\n"); return 2; }
  const int m = std::atoi(argv[2]), n = std::atoi(argv[3]), l = std::atoi(argv[4]);
  const std::string out = argv[5];
  Mat_m A(m, n);
  { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(A.data()), sizeof(double) * m * n); if (!f) return 3; }

  // tests/rSVD_test.cpp:65-72 -- the caller pre-sizes U, S, V (V even as l x n); rSVD assigns them
  Mat_m U = Mat_m::Zero(m, l); Vec_v S = Vec_v::Zero(l); Mat_m V = Mat_m::Zero(l, n);
  Mat_m Omega = generateOmega(n, l);
  rSVD(A, U, S, V, l, SVDMethod::Jacobi, Omega, 2);
  dump(out + "_Omega.bin", Omega.data(), (size_t)n * l);
  dump(out + "_U.bin", U.data(), (size_t)U.rows() * U.cols()); dump(out + "_S.bin", S.data(), (size_t)S.size()); dump(out + "_V.bin", V.data(), (size_t)V.rows() * V.cols());
  std::printf("rSVD: U %ld x %ld, S %ld, V %ld x %ld\n", (long)U.rows(), (long)U.cols(), (long)S.size(), (long)V.rows(), (long)V.cols());

  // the reference's 6-argument call (Omega drawn internally, q = 2)
  Mat_m U2, V2; Vec_v S2;
  rSVD(A, U2, S2, V2, l, SVDMethod::ParallelJacobi);
  dump(out + "_S2.bin", S2.data(), (size_t)S2.size());

  Mat_m Q = Mat_m::Zero(m, l);
  intermediate_step(A, Q, Omega, l, 2);
  dump(out + "_Q.bin", Q.data(), (size_t)m * l);

  // tests/svd_test.cpp:58-67
  SVD<SVDMethod::Jacobi> svd(A); svd.compute();
  Vec_v Sf = svd.getS(); dump(out + "_Sfull.bin", Sf.data(), (size_t)Sf.size());
  std::printf("SVD<Jacobi>: U %ld x %ld, V %ld x %ld\n", (long)svd.getU().rows(), (long)svd.getU().cols(), (long)svd.getV().rows(), (long)svd.getV().cols());

  // tests/QRTest.cpp:64
  if (m >= n) {
    Mat_m Qr, Rr; qr_decomposition_reduced(A, Qr, Rr);
    dump(out + "_QRq.bin", Qr.data(), (size_t)m * n); dump(out + "_QRr.bin", Rr.data(), (size_t)n * n);
    QRReducedDecomposition<double> qr(A); Mat_m Qc, Rc; qr.decompose(Qc, Rc);
    double d = 0; for (std::ptrdiff_t i = 0; i < Rr.size(); ++i) d += std::abs(Rr.data()[i] - Rc.data()[i]);
    std::printf("QR class vs free function |dR|_1 = %g\n", d);
  }
  Mat_m B;   // PM ignores it
  double sigma; Vec_v u, v; PM(A, B, sigma, u, v);
  dump(out + "_pm.bin", &sigma, 1);

  Mat C = manualMatrixMultiply(A, Omega);
  dump(out + "_AOmega.bin", C.data(), (size_t)m * l);
  int threw = 0;
  try { manualMatrixMultiply(A, A.rows() == A.cols() ? Omega : A); } catch (const std::invalid_argument&) { threw = 1; }
  try { rSVD(A, U, S, V, l, static_cast<SVDMethod>(7)); } catch (const std::invalid_argument&) { threw += 2; }
  std::printf("invalid_argument paths: %d\n", threw);

  // src/JacobiOperations.cpp:6-103 -- host-side rotation helpers keep the reference's arithmetic
  {
    Mat_m M(3, 3); double vals[9] = {4, 1, 2, 0.5, 3, 1, 0.25, 0.75, 2};
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) M(i, j) = vals[i + 3 * j];
    double cl, sl, cr, sr;
    const bool real = svd_precondition_2x2_block_to_be_real(M, 1, 0, 4.0);
    real_2x2_jacobi_svd(M, cl, sl, cr, sr, 1, 0);
    applyOnTheLeft(M, 1, 0, cl, sl); applyOnTheRight(M, 1, 0, cr, sr);
    std::printf("rot2x2 real=%d offdiag=%.3e %.3e cl=%.17g sl=%.17g cr=%.17g sr=%.17g\n", (int)real, M(1, 0), M(0, 1), cl, sl, cr, sr);
    double cl2, sl2, cr2, sr2; Mat_m M2(2, 2); M2(0, 0) = 1; M2(0, 1) = 1e-12; M2(1, 0) = 0; M2(1, 1) = 2;
    real_2x2_jacobi_svd_par(M2, cl2, sl2, cr2, sr2, 0, 1);
    std::printf("rot2x2_par cr=%.17g sr=%.17g\n", cr2, sr2);
    JacobiRotation r1(0.6, 0.8); Vec_v e(2); e(0) = 1; e(1) = 0; Vec_v re = r1.apply(e);
    std::printf("rotation apply %.17g %.17g\n", re(0), re(1));
  }
  JacobiRotation rot; bool okj = rot.makeJacobi(2.0, 0.5, 1.0);
  std::printf("makeJacobi ok=%d c=%.17g s=%.17g\n", (int)okj, rot.getC(), rot.getS());
  return 0;
}
