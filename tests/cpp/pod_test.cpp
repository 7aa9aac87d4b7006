// C++ caller of the drop-in include/POD.hpp, shaped like the reference's use in POD/ParametricDiffusion1D/src/Diff1D.cpp
// (build the snapshot matrix, construct POD(S, r, tol, svd_type), read pod.W / pod.sigma).
//   usage: pod_test <S.bin> <Nh> <ns> <r> <tol> <svd_type> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "POD.hpp"

static void dump(const std::string& path, const double* p, size_t n) {
  std::ofstream f(path, std::ios::binary); f.write(reinterpret_cast<const char*>(p), sizeof(double) * n);
}

int main(int argc, char** argv) {
  if (argc < 8) { std::fprintf(stderr, "usage\n"); return 2; }
  const int Nh = std::atoi(argv[2]), ns = std::atoi(argv[3]), r = std::atoi(argv[4]), svd_type = std::atoi(argv[6]);
  const double tol = std::atof(argv[5]);
  const std::string out = argv[7];
  Mat_m S(Nh, ns);
  { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(S.data()), sizeof(double) * Nh * ns); if (!f) return 3; }
  POD naive(S, r, svd_type);
  POD standard(S, r, tol, svd_type);
  std::printf("naive W %ld x %ld sigma %ld\n", (long)naive.W.rows(), (long)naive.W.cols(), (long)naive.sigma.size());
  std::printf("standard W %ld x %ld sigma %ld\n", (long)standard.W.rows(), (long)standard.W.cols(), (long)standard.sigma.size());
  dump(out + "_naive_W.bin", naive.W.data(), (size_t)naive.W.size()); dump(out + "_naive_sigma.bin", naive.sigma.data(), (size_t)naive.sigma.size());
  dump(out + "_std_W.bin", standard.W.data(), (size_t)standard.W.size()); dump(out + "_std_sigma.bin", standard.sigma.data(), (size_t)standard.sigma.size());
  Mat_m Xh = Mat_m::Zero(Nh, Nh);            // SPD tridiagonal operator (the same one tests/golden/make_golden.py builds)
  for (int i = 0; i < Nh; ++i) { Xh(i, i) = 2.0; if (i + 1 < Nh) { Xh(i, i + 1) = -0.5; Xh(i + 1, i) = -0.5; } }
  Mat_m D = Mat_m::Identity(ns, ns);
  for (int i = 0; i < ns; ++i) D(i, i) = 1.0 + 0.5 * (i % 3);
  POD energy(S, Xh, r, tol, svd_type), weight(S, Xh, D, r, tol, svd_type);
  std::printf("energy W %ld x %ld, weight W %ld x %ld\n", (long)energy.W.rows(), (long)energy.W.cols(), (long)weight.W.rows(), (long)weight.W.cols());
  dump(out + "_energy_W.bin", energy.W.data(), (size_t)energy.W.size()); dump(out + "_weight_W.bin", weight.W.data(), (size_t)weight.W.size());
  dump(out + "_energy_sigma.bin", energy.sigma.data(), (size_t)energy.sigma.size()); dump(out + "_weight_sigma.bin", weight.sigma.data(), (size_t)weight.sigma.size());
  return 0;
}
