// Caller of the OLDER API (reference image_compression/tests/rSVD_test1.cpp:69-76: k = 5, p = 10, l = k + p; SVD_test2.cpp:30-44)
// against include/image_compression/*.hpp.   usage: rsvd_v1_test <in.bin> <m> <n> <l> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include "image_compression/rSVD.hpp"
#include "image_compression/QR.hpp"

static void dump(const std::string& path, const double* p, size_t n) { std::ofstream f(path, std::ios::binary); f.write(reinterpret_cast<const char*>(p), sizeof(double) * n); }

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int m = std::atoi(argv[2]), n = std::atoi(argv[3]), l = std::atoi(argv[4]); const std::string out = argv[5];
  Mat_m A(m, n);
  { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(A.data()), sizeof(double) * m * n); if (!f) return 3; }
  Mat_m U = Mat_m::Zero(m, l), V = Mat_m::Zero(n, l); Vec_v S = Vec_v::Zero(l);
  rSVD(A, U, S, V, l);
  std::printf("rSVD(5 args): U %ld x %ld, S %ld, V %ld x %ld\n", (long)U.rows(), (long)U.cols(), (long)S.size(), (long)V.rows(), (long)V.cols());
  dump(out + "_U.bin", U.data(), (size_t)m * l); dump(out + "_S.bin", S.data(), (size_t)l); dump(out + "_V.bin", V.data(), (size_t)n * l);
  // singularValueDecomposition on a copy: dim = 4 triplets, A deflated in place
  Mat_m A2 = A; const int dim = 4;
  Vec_v s2 = Vec_v::Zero(dim); Mat_m U2 = Mat_m::Zero(m, dim), V2;
  singularValueDecomposition(A2, s2, U2, V2, dim);
  std::printf("SVD dim=4: V %ld x %ld\n", (long)V2.rows(), (long)V2.cols());
  dump(out + "_s2.bin", s2.data(), dim); dump(out + "_U2.bin", U2.data(), (size_t)m * dim); dump(out + "_V2.bin", V2.data(), (size_t)n * dim);
  dump(out + "_A2.bin", A2.data(), (size_t)m * n);
  Mat_m B; double sg; Vec_v u, v; powerMethod(A, B, sg, u, v);
  dump(out + "_pm.bin", &sg, 1);
  return 0;
}
