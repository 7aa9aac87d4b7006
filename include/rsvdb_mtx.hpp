// MatrixMarket IO for callers of the drop-in headers: the reference's test mains read their inputs with
// `Eigen::loadMarket` into a sparse matrix and densify it (tests/rSVD_test.cpp:54-57), and write U, S, V with
// `Eigen::saveMarket` (:113-115).  With Eigen those calls keep working; without it these two functions do the same job
// ("coordinate real general", 1-based, duplicate entries add up; "array real general" on output).
#ifndef RSVDB_MTX_HPP
#define RSVDB_MTX_HPP

#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

#include "rsvdb_dense.hpp"

namespace rsvdb {

inline bool load_market_dense(Mat_m& A, const std::string& path) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  if (!std::getline(f, line) || line.rfind("%%MatrixMarket", 0) != 0) return false;
  const bool array = line.find("array") != std::string::npos;
  const bool symmetric = line.find("symmetric") != std::string::npos;
  const bool pattern = line.find("pattern") != std::string::npos;
  while (std::getline(f, line)) if (!line.empty() && line[0] != '%') break;
  std::istringstream hdr(line);
  long m = 0, n = 0, nnz = 0;
  hdr >> m >> n; if (!array) hdr >> nnz;
  A = Mat_m::Zero(m, n);
  if (array) {
    for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) { double v; if (!(f >> v)) return false; A(i, j) = v; }
    return true;
  }
  for (long e = 0; e < nnz; ++e) {
    long i, j; double v = 1.0;
    if (!(f >> i >> j)) return false;
    if (!pattern && !(f >> v)) return false;
    A(i - 1, j - 1) += v;
    if (symmetric && i != j) A(j - 1, i - 1) += v;
  }
  return true;
}

template <typename M>
inline bool save_market_dense(const M& A, const std::string& path) {
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) return false;
  const long m = static_cast<long>(A.rows()), n = static_cast<long>(A.cols());
  std::fprintf(f, "%%%%MatrixMarket matrix array real general\n%ld %ld\n", m, n);
  for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) std::fprintf(f, "%.17g\n", A.data()[i + j * m]);
  std::fclose(f);
  return true;
}

}  // namespace rsvdb
#endif
