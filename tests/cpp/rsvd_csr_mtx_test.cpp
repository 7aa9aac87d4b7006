// The reference's tests/rSVD_test.cpp flow on a SPARSE MatrixMarket input without densifying it (north_star: "a vectorised,
// coalesced CSR SpMM ... for sparse .mtx inputs"): load_market_csr -> rSVD(CsrMatrix, U, S, V, l) -> ||A - U S V^T||_F evaluated
// against the densified matrix, U / S / V written like the reference's main does.
//   usage: rsvd_csr_mtx_test <file.mtx> <l> <out_prefix>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rSVD.hpp"
#include "rsvdb_mtx.hpp"

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage\n"); return 2; }
  const int l = std::atoi(argv[2]);
  rsvdb::CsrMatrix A;
  if (!rsvdb::load_market_csr(A, argv[1])) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
  Mat_m Ad;
  if (!rsvdb::load_market_dense(Ad, argv[1])) return 3;               // only for the error norm below (the reference densifies, :57)
  Mat_m U, V; Vec_v S;
  rsvdb::rSVD(A, U, S, V, l);
  double err2 = 0.0, nrm2 = 0.0;
  for (std::ptrdiff_t j = 0; j < Ad.cols(); ++j)
    for (std::ptrdiff_t i = 0; i < Ad.rows(); ++i) {
      double r = Ad(i, j);
      for (std::ptrdiff_t k = 0; k < S.size(); ++k) r -= U(i, k) * S(k) * V(j, k);
      err2 += r * r; nrm2 += Ad(i, j) * Ad(i, j);
    }
  std::printf("Size: %ld, %ld  nnz: %ld\nnorm of diff : %.10g\nnorm of A : %.10g\n", (long)A.rows, (long)A.cols, (long)A.nnz(), std::sqrt(err2), std::sqrt(nrm2));
  const std::string out = argv[3];
  rsvdb::save_market_dense(S, out + "_S.mtx"); rsvdb::save_market_dense(U, out + "_U.mtx"); rsvdb::save_market_dense(V, out + "_V.mtx");
  bool threw = false;
  try { rsvdb::rSVD(A, U, S, V, l, 7); } catch (const std::invalid_argument&) { threw = true; }      // src/rSVD.cpp:122-123
  std::printf("invalid method throws: %d\n", threw ? 1 : 0);
  return 0;
}
