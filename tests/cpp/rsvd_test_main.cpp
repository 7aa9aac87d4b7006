// The reference's main test driver, tests/rSVD_test.cpp, restated against the drop-in headers: for every .mtx file of an
// input directory, densify, run rSVD(A, U, S, V, l = 16, SVDMethod::Jacobi), print size / time / ||A - U S V^T||_F and
// write <name>_{U,S,V}.mtx.  (k = 0, p = 16: tests/rSVD_test.cpp:65-67.)   usage: rsvd_test_main <input_dir> <output_dir>
#include <chrono>
#include <cmath>
#include <filesystem>
#include <iostream>
#include <vector>

#include "rSVD.hpp"
#include "rsvdb_mtx.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { std::cerr << "usage: rsvd_test_main <input_dir> <output_dir>\n"; return 2; }
  std::cout << "test rSVD reduced" << std::endl;
  const std::filesystem::path inputDir = argv[1], outputDir = argv[2];
  std::filesystem::create_directories(outputDir);
  std::vector<std::string> fileNames;
  for (const auto& entry : std::filesystem::directory_iterator(inputDir))
    if (entry.is_regular_file() && entry.path().extension() == ".mtx") fileNames.push_back(entry.path().filename().string());
  for (auto& fileName : fileNames) {
    Mat_m A;
    if (!rsvdb::load_market_dense(A, (inputDir / fileName).string())) { std::cerr << "cannot read " << fileName << "\n"; return 3; }
    const auto start = std::chrono::high_resolution_clock::now();
    const int m = static_cast<int>(A.rows()), n = static_cast<int>(A.cols());
    const int k = 0, p = 16, l = k + p;
    Mat_m U = Mat_m::Zero(m, l); Vec_v S = Vec_v::Zero(l); Mat_m V = Mat_m::Zero(l, n);
    rSVD(A, U, S, V, l, SVDMethod::Jacobi);
    const auto end = std::chrono::high_resolution_clock::now();
    // ||A - U diag(S) V^T||_F, computed by the caller as in tests/rSVD_test.cpp:77-84
    double nd = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double r = 0.0;
        for (int c = 0; c < static_cast<int>(S.size()); ++c) r += U(i, c) * S(c) * V(j, c);
        const double d = A(i, j) - r; nd += d * d;
      }
    std::cout << "\nDataset: " << fileName << "\nSize: " << m << ", " << n << "\nNumber of Processors: 1\nExecution time: "
              << std::chrono::duration<double>(end - start).count() << " seconds\nnorm of diff : " << std::sqrt(nd)
              << "\n-------------------------\n" << std::endl;
    const std::string stem = fileName.substr(0, fileName.find_last_of('.'));
    rsvdb::save_market_dense(S, (outputDir / (stem + "_S.mtx")).string());
    rsvdb::save_market_dense(U, (outputDir / (stem + "_U.mtx")).string());
    rsvdb::save_market_dense(V, (outputDir / (stem + "_V.mtx")).string());
  }
  return 0;
}
