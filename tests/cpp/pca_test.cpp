// C++ caller in the shape of reference PCA/tests/pca_test.cpp:60-91 (load, PCA<ParallelJacobi>(data, normalize), summary(),
// saveResults()), compiled against the drop-in include/PCA_class.hpp and linked to librsvdb.so.  Reads a column-major
// binary matrix instead of the tourists table; dumps the quantities the Python side compares with the oracle.
//   usage: pca_test <in.bin> <m> <n> <yes|no> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "PCA_class.hpp"

static void dump(const std::string& path, const double* p, size_t n) {
  std::ofstream f(path, std::ios::binary); f.write(reinterpret_cast<const char*>(p), sizeof(double) * n);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage\n"); return 2; }
  const int m = std::atoi(argv[2]), n = std::atoi(argv[3]);
  const bool normalize = std::string(argv[4]) == "yes";
  const std::string out = argv[5];
  Mat_m data(m, n);
  { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(data.data()), sizeof(double) * m * n); if (!f) return 3; }
  try {
    PCA<SVDMethod::ParallelJacobi> pca(data, normalize);
    pca.summary();
    pca.saveResults(out + "_results.txt");
    const Vec_v ev = pca.explainedVariance(), ratio = pca.explainedVarianceRatio();
    dump(out + "_ev.bin", ev.data(), (size_t)ev.size()); dump(out + "_ratio.bin", ratio.data(), (size_t)ratio.size());
    const Mat_m sc = pca.scores(), ld = pca.loadings();
    dump(out + "_scores.bin", sc.data(), (size_t)sc.size()); dump(out + "_loadings.bin", ld.data(), (size_t)ld.size());
    const Mat_m pr = pca.projectToPCA(data), rc = pca.reconstructFromPCA(pr);
    dump(out + "_project.bin", pr.data(), (size_t)pr.size()); dump(out + "_reconstruct.bin", rc.data(), (size_t)rc.size());
    std::printf("orthogonality %.3e\n", pca.checkOrthogonality());
    // setNormalization re-runs initialize() (PCA_class.hpp:69-72)
    pca.setNormalization(!normalize);
    const Vec_v ev2 = pca.explainedVariance(); dump(out + "_ev_flipped.bin", ev2.data(), (size_t)ev2.size());
    // addData appends rows and re-initialises (:57-61)
    pca.addData(data);
    std::printf("after addData: scores %ld x %ld\n", (long)pca.scores().rows(), (long)pca.scores().cols());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
  // assertDataValid (:50-54)
  try { Mat_m tiny(1, 5); PCA<SVDMethod::Jacobi> bad(tiny); std::printf("no throw\n"); }
  catch (const std::invalid_argument& e) { std::printf("invalid_argument: %s\n", e.what()); }
  return 0;
}
