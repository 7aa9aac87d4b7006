// The flow of reference image_compression/main/main.cpp:44-80 (load -> downscale -> normalize -> compress -> deNormalize ->
// upscale -> save / ratio) against the drop-in include/image_compression/image_comp.hpp, with a raw matrix instead of a PNG.
//   usage: image_test <in.bin> <m> <n> <k> <out_prefix>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "image_compression/image_comp.hpp"

static void dump(const std::string& path, const double* p, size_t n) {
  std::ofstream f(path, std::ios::binary); f.write(reinterpret_cast<const char*>(p), sizeof(double) * n);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage\n"); return 2; }
  const int m = std::atoi(argv[2]), n = std::atoi(argv[3]), k = std::atoi(argv[4]);
  const std::string out = argv[5];
  Mat_m A(m, n);
  { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(A.data()), sizeof(double) * m * n); if (!f) return 3; }
  Image img;
  img.setMatrix(A);
  img.downscale(2);
  img.normalize();
  dump(out + "_normalized.bin", img.getMatrix().data(), (size_t)img.getMatrix().size());
  std::printf("range %.17g %.17g\n", img.getMin(), img.getMax());
  img.compress(k);
  Mat_m rec = img.reconstruct();
  dump(out + "_rec.bin", rec.data(), (size_t)rec.size());
  dump(out + "_S.bin", img.getS().data(), (size_t)img.getS().size());
  Mat_m rec2 = img.reconstruct_denormalized();
  dump(out + "_rec_denorm.bin", rec2.data(), (size_t)rec2.size());
  img.compress_parallel(k);                    // overwrites image_matrix with the reconstruction
  img.deNormalize();
  img.upscale(2);
  dump(out + "_final.bin", img.getMatrix().data(), (size_t)img.getMatrix().size());
  std::printf("final %ld x %ld ratio %.6f\n", (long)img.getMatrix().rows(), (long)img.getMatrix().cols(), img.get_compression_ratio());
  img.save_compressed(out + "_compressed.bin");
  Image img2; img2.load_compressed(out + "_compressed.bin");
  std::printf("compressed file: U %ld x %ld\n", (long)img2.getU().rows(), (long)img2.getU().cols());
  return 0;
}
