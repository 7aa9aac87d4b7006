// Drop-in for reference PCA/include/PCA_class.hpp: template<SVDMethod> class PCA : public SVD<method>, same members.
// initialize() is ONE call into librsvdb.so (rsvdb_pca_host: upload once, column means / centring / optional stddev
// scaling and SVD<method> on the device); projectToPCA / reconstructFromPCA are device GEMMs with the mean folded in as
// a rank-1 term.  The O(k) / O(m k) getters (explained variance, ratio, scores = U diag(S)) are plain host loops over
// the returned factors, exactly the reference's expressions.
#ifndef CLASS_PCA_HPP
#define CLASS_PCA_HPP

#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>

#include "SVD_class.hpp"

template <SVDMethod method>
class PCA : public SVD<method> {
 public:
  using Mat = Mat_m;
  using Vec = Vec_v;

  // reference :18-22
  PCA(const Mat& data, bool normalize = false) : SVD<method>(data), data_(data), normalize_(normalize) { initialize(); }

  // reference :24-47 (the progress prints are dropped)
  void initialize() {
    assertDataValid();
    const std::ptrdiff_t m = data_.rows(), n = data_.cols(), k = m < n ? m : n;
    Mat U(m, k), V(n, k); Vec S(k);
    mean_ = Vec::Zero(n); stddev_ = Vec::Zero(n);
    int found = 0;
    rsvdb::check(rsvdb::default_context(),
                 rsvdb_pca_host(rsvdb::default_context(), data_.data(), m, n, m, normalize_ ? 1 : 0, static_cast<int>(method), 0, 0,
                                mean_.data(), stddev_.data(), U.data(), m, S.data(), V.data(), n, &found));
    SVD<method>::setResults(U, S, V);
  }

  // reference :50-54
  void assertDataValid() const {
    if (data_.rows() < 2 || data_.cols() < 2) throw std::invalid_argument("PCA requires at least 2 rows and 2 columns.");
  }

  // reference :57-61
  void addData(const Mat& newData) {
    const std::ptrdiff_t m = data_.rows(), n = data_.cols(), a = newData.rows();
    Mat d(m + a, n);
    for (std::ptrdiff_t j = 0; j < n; ++j) {
      for (std::ptrdiff_t i = 0; i < m; ++i) d(i, j) = data_(i, j);
      for (std::ptrdiff_t i = 0; i < a; ++i) d(m + i, j) = newData(i, j);
    }
    data_ = d;
    initialize();
  }

  // reference :63-66 (scales the stored data by the UNcentred second moment)
  void normalizeData() {
    const std::ptrdiff_t m = data_.rows(), n = data_.cols();
    stddev_ = Vec::Zero(n);
    for (std::ptrdiff_t j = 0; j < n; ++j) {
      double s = 0.0;
      for (std::ptrdiff_t i = 0; i < m; ++i) s += data_(i, j) * data_(i, j);
      stddev_(j) = std::sqrt(s / static_cast<double>(m - 1));
      for (std::ptrdiff_t i = 0; i < m; ++i) data_(i, j) /= stddev_(j);
    }
  }

  // reference :69-72
  void setNormalization(bool normalize) { normalize_ = normalize; initialize(); }

  // reference :75-78
  Vec explainedVariance() const {
    Vec s = SVD<method>::getS();
    const double d = std::sqrt(static_cast<double>(data_.rows() - 1));
    for (std::ptrdiff_t i = 0; i < s.size(); ++i) s(i) /= d;
    return s;
  }
  // reference :80-83
  Vec explainedVarianceRatio() const {
    Vec v = explainedVariance();
    const double d = static_cast<double>(data_.rows() - 1);
    double tot = 0.0;
    for (std::ptrdiff_t i = 0; i < v.size(); ++i) tot += v(i) * v(i);
    for (std::ptrdiff_t i = 0; i < v.size(); ++i) v(i) = (v(i) * v(i) / d) / (tot / d);
    return v;
  }
  // reference :85-87
  Mat scores() const {
    Mat u = SVD<method>::getU(); const Vec s = SVD<method>::getS();
    for (std::ptrdiff_t j = 0; j < u.cols(); ++j) for (std::ptrdiff_t i = 0; i < u.rows(); ++i) u(i, j) *= s(j);
    return u;
  }
  // reference :89-91
  Mat loadings() const { return SVD<method>::getV(); }

  // reference :93-95
  Mat projectToPCA(const Mat& data) {
    const Mat V = SVD<method>::getV();
    Mat out(data.rows(), V.cols());
    rsvdb::check(rsvdb::default_context(),
                 rsvdb_pca_project_host(rsvdb::default_context(), data.data(), data.rows(), data.cols(), data.rows(), mean_.data(), V.data(),
                                        V.rows(), static_cast<int>(V.cols()), out.data(), data.rows()));
    return out;
  }
  // reference :97-99
  Mat reconstructFromPCA(const Mat& pcData) {
    const Mat V = SVD<method>::getV();
    Mat out(pcData.rows(), V.rows());
    rsvdb::check(rsvdb::default_context(),
                 rsvdb_pca_reconstruct_host(rsvdb::default_context(), pcData.data(), pcData.rows(), static_cast<int>(pcData.cols()),
                                            pcData.rows(), mean_.data(), V.data(), V.rows(), V.rows(), out.data(), pcData.rows()));
    return out;
  }

  // reference :101-144 (same file layout)
  void saveResults(const std::string& filename) {
    std::ofstream outFile(filename);
    Vec cum = explainedVarianceRatio();
    for (std::ptrdiff_t i = 1; i < cum.size(); ++i) cum(i) += cum(i - 1);
    outFile << "\nCumulative Explained Variance:\n";
    for (std::ptrdiff_t i = 0; i < cum.size(); ++i) outFile << cum(i) << std::endl;
    const Mat sc = scores();
    outFile << "\nScores:\n";
    for (std::ptrdiff_t i = 0; i < sc.rows(); ++i) {
      for (std::ptrdiff_t j = 0; j < sc.cols(); ++j) { outFile << sc(i, j); if (j < sc.cols() - 1) outFile << ", "; }
      outFile << std::endl;
    }
    const Mat ld = loadings();
    outFile << "\nLoadings:\n";
    for (std::ptrdiff_t i = 0; i < ld.rows(); ++i) {
      for (std::ptrdiff_t j = 0; j < ld.cols(); ++j) { outFile << ld(i, j); if (j < ld.cols() - 1) outFile << ", "; }
      outFile << std::endl;
    }
    outFile.close();
  }

  // reference :147-151 (a k x k Gram matrix of the returned loadings, on the host)
  double checkOrthogonality() const {
    const Mat V = SVD<method>::getV();
    double s = 0.0;
    for (std::ptrdiff_t a = 0; a < V.cols(); ++a)
      for (std::ptrdiff_t b = 0; b < V.cols(); ++b) {
        double d = 0.0;
        for (std::ptrdiff_t i = 0; i < V.rows(); ++i) d += V(i, a) * V(i, b);
        d -= (a == b) ? 1.0 : 0.0;
        s += d * d;
      }
    return std::sqrt(s);
  }

  // reference :153-196
  void summary() const {
    const Vec exp_var = explainedVariance();
    const int numComponents = static_cast<int>(exp_var.size());
    const Vec proportion = explainedVarianceRatio();
    Vec cumulative = proportion;
    for (int i = 1; i < numComponents; ++i) cumulative(i) += cumulative(i - 1);
    std::cout << std::fixed << std::setprecision(6);
    std::cout << "Importance of components:\n";
    std::cout << std::setw(25) << std::left << "Component";
    for (int i = 1; i <= numComponents; ++i) std::cout << std::setw(15) << std::left << ("Comp." + std::to_string(i));
    std::cout << std::endl;
    const char* names[3] = {"Standard deviation", "Proportion of Variance", "Cumulative Proportion"};
    const Vec* rows[3] = {&exp_var, &proportion, &cumulative};
    for (int r = 0; r < 3; ++r) {
      std::cout << std::setw(25) << std::left << names[r];
      for (int i = 0; i < numComponents; ++i) std::cout << std::setw(15) << std::left << (*rows[r])(i);
      std::cout << std::endl;
    }
  }

  // additive: mean_ / stddev_ have no getters in the reference
  Vec mean() const { return mean_; }
  Vec stddev() const { return stddev_; }

 private:
  Mat data_;
  Vec mean_;
  Vec stddev_;
  bool normalize_;
};

#endif  // CLASS_PCA_HPP
