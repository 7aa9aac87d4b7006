// MatrixMarket IO for callers of the drop-in headers: the reference's test mains read their inputs with
// `Eigen::loadMarket` into a sparse matrix and densify it (tests/rSVD_test.cpp:54-57), and write U, S, V with
// `Eigen::saveMarket` (:113-115).  With Eigen those calls keep working; without it these two functions do the same job
// ("coordinate real general", 1-based, duplicate entries add up; "array real general" on output).
#ifndef RSVDB_MTX_HPP
#define RSVDB_MTX_HPP

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

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

// A MatrixMarket coordinate file as CSR (int64 row pointers, int32 column indices, rows sorted by column, duplicates kept
// as separate entries -- they add up in the products, like Eigen's densification adds them up).  The reference densifies
// every .mtx before rSVD (tests/rSVD_test.cpp:54-57); this is the entry to the path that does not.
struct CsrMatrix {
  std::int64_t rows = 0, cols = 0;
  std::vector<std::int64_t> rowptr;
  std::vector<std::int32_t> colidx;
  std::vector<double> values;
  std::int64_t nnz() const { return static_cast<std::int64_t>(values.size()); }
};

inline bool load_market_csr(CsrMatrix& A, const std::string& path) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line;
  if (!std::getline(f, line) || line.rfind("%%MatrixMarket", 0) != 0 || line.find("coordinate") == std::string::npos) return false;
  const bool symmetric = line.find("symmetric") != std::string::npos;
  const bool pattern = line.find("pattern") != std::string::npos;
  while (std::getline(f, line)) if (!line.empty() && line[0] != '%') break;
  std::istringstream hdr(line);
  long long m = 0, n = 0, nnz = 0;
  hdr >> m >> n >> nnz;
  if (m <= 0 || n <= 0 || nnz < 0 || n > 2147483647LL) return false;
  std::vector<std::int64_t> ri; std::vector<std::int32_t> ci; std::vector<double> vv;
  ri.reserve(static_cast<size_t>(nnz)); ci.reserve(static_cast<size_t>(nnz)); vv.reserve(static_cast<size_t>(nnz));
  for (long long e = 0; e < nnz; ++e) {
    long long i, j; double v = 1.0;
    if (!(f >> i >> j)) return false;
    if (!pattern && !(f >> v)) return false;
    if (i < 1 || i > m || j < 1 || j > n) return false;
    ri.push_back(i - 1); ci.push_back(static_cast<std::int32_t>(j - 1)); vv.push_back(v);
    if (symmetric && i != j) { ri.push_back(j - 1); ci.push_back(static_cast<std::int32_t>(i - 1)); vv.push_back(v); }
  }
  std::vector<size_t> order(ri.size());
  std::iota(order.begin(), order.end(), size_t(0));
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ri[a] != ri[b] ? ri[a] < ri[b] : ci[a] < ci[b]; });
  A.rows = m; A.cols = n;
  A.rowptr.assign(static_cast<size_t>(m) + 1, 0);
  A.colidx.resize(order.size()); A.values.resize(order.size());
  for (size_t k = 0; k < order.size(); ++k) { A.colidx[k] = ci[order[k]]; A.values[k] = vv[order[k]]; ++A.rowptr[static_cast<size_t>(ri[order[k]]) + 1]; }
  for (size_t i = 0; i < static_cast<size_t>(m); ++i) A.rowptr[i + 1] += A.rowptr[i];
  return true;
}

// rSVD of a sparse matrix without densifying it: the same argument meaning as rSVD(A, U, S, V, l, method) (include/rSVD.hpp:14,
// src/rSVD.cpp:72-133; q = 2 there), outputs assigned like the dense overload.  method as int (enum class SVDMethod's values).
inline void rSVD(const CsrMatrix& A, Mat_m& U, Vec_v& S, Mat_m& V, int l, int method = 0, int q = 2, std::uint64_t seed = 0x5eedULL,
                 const Mat_m* Omega = nullptr) {
  rsvdb_ctx* c = default_context();
  if (method != 0 && method != 1 && method != 2) throw std::invalid_argument("Unsupported SVD method");
  if (Omega && (Omega->rows() != A.cols || Omega->cols() != l)) throw std::invalid_argument("Omega must be n x l");
  const std::ptrdiff_t m = static_cast<std::ptrdiff_t>(A.rows), n = static_cast<std::ptrdiff_t>(A.cols), k = l < n ? l : n;
  Mat_m Un(m, k), Vn(n, k); Vec_v Sn(k);
  check(c, rsvdb_rsvd_csr_host(c, A.rows, A.cols, A.nnz(), A.rowptr.data(), A.colidx.data(), A.values.data(), Omega ? Omega->data() : nullptr,
                               Omega ? Omega->rows() : 0, seed, l, q, method, Un.data(), m, Sn.data(), Vn.data(), n));
  U = Un; S = Sn;
  if (method == 1) {                           // Power back-end: V is n x n with the vectors in rows (include/SVD_class.hpp:83,214)
    Mat_m Vr = Mat_m::Identity(n, n);
    for (std::ptrdiff_t i = 0; i < k; ++i) for (std::ptrdiff_t j = 0; j < n; ++j) Vr(i, j) = Vn(j, i);
    V = Vr;
  } else V = Vn;
}

}  // namespace rsvdb
#endif
