// Drop-in for reference image_compression/include/image_comp.hpp: class Image with the same members, minus the stb codec
// (load / save of PNG/JPEG, image_com.cpp:18-64), which is outside the accelerated path -- the pixel matrix is handed in
// with setMatrix() instead.  normalize / deNormalize / compress / reconstruct run in librsvdb.so on the GPU;
// downscale / upscale / save_compressed / load_compressed / get_compression_ratio are index shuffles and file IO and stay
// on the host with the reference's exact semantics (including the lossy one-byte-per-entry compressed file format).
#ifndef IMAGE_COMP_HPP
#define IMAGE_COMP_HPP

#include <fstream>
#include <iostream>
#include <memory>
#include <string>

#include "rSVD.hpp"   // include/image_compression/rSVD.hpp (older 5-argument API)

class Image {
 private:
  Mat_m image_matrix;
  Mat_m left_singular; Vec_v singular; Mat_m right_singular;
  int originalWidth = 0, originalHeight = 0, channels = 1;
  double original_min = 0.0, original_max = 0.0;
  int degree = 0;

 public:
  Image() {}
  Image(int width, int height) { originalWidth = width; originalHeight = height; }

  // additive: stands in for load() (image_com.cpp:18-44 stores the TRANSPOSE of the decoded height x width picture)
  void setMatrix(const Mat_m& m) { image_matrix = m; originalWidth = static_cast<int>(m.rows()); originalHeight = static_cast<int>(m.cols()); }
  const Mat_m& getMatrix() const { return image_matrix; }
  const Mat_m& getU() const { return left_singular; }
  const Vec_v& getS() const { return singular; }
  const Mat_m& getV() const { return right_singular; }
  double getMin() const { return original_min; }
  double getMax() const { return original_max; }

  // image_com.cpp:251-264
  void normalize() {
    rsvdb_ctx* c = rsvdb::default_context();
    rsvdb::check(c, rsvdb_image_normalize_host(c, image_matrix.data(), image_matrix.rows(), image_matrix.cols(), image_matrix.rows(), 0,
                                               &original_min, &original_max));
    if (!(original_min < original_max)) std::cerr << "Warning: Unable to normalize image. Empty or undefined pixel value range." << std::endl;
  }
  // image_com.cpp:270-281
  void deNormalize() {
    if (!(original_min < original_max)) { std::cerr << "Warning: Unable to deNormalize image. Empty or undefined pixel value range." << std::endl; return; }
    rsvdb_ctx* c = rsvdb::default_context();
    rsvdb::check(c, rsvdb_image_normalize_host(c, image_matrix.data(), image_matrix.rows(), image_matrix.cols(), image_matrix.rows(), 1,
                                               &original_min, &original_max));
  }
  // image_com.cpp:288-317
  void compress(int k = -1) { compress_impl(k, 0); }
  // additive: normalize() + compress(k) with one upload; image_matrix itself is left as it was
  void normalize_and_compress(int k = -1) { compress_impl(k, 1); }
  // image_com.cpp:325-404 splits the image over a square MPI process grid and compresses the blocks independently; on one
  // GPU the whole image is one block (the numProcesses == 1 case), after which the reference overwrites image_matrix with
  // the reconstruction (:357-372)
  void compress_parallel(int k = -1) { compress_impl(k, 0); image_matrix = reconstruct(); }

  // image_com.cpp:184-190
  Mat_m reconstruct() { return reconstruct_impl(0); }
  Mat_m reconstruct_denormalized() { return reconstruct_impl(1); }   // additive: reconstruct() + deNormalize() in one pass

  // image_com.cpp:193-217
  void downscale(int scale_factor = -1) {
    if (scale_factor == -1) scale_factor = 2;
    const int nw = originalWidth / scale_factor, nh = originalHeight / scale_factor;
    Mat_m d(nh, nw);
    for (int i = 0; i < nh; ++i) for (int j = 0; j < nw; ++j) d(i, j) = image_matrix(i * scale_factor, j * scale_factor);
    image_matrix = d; originalWidth = nw; originalHeight = nh;
  }
  // image_com.cpp:219-244
  void upscale(int scale_factor = -1) {
    if (scale_factor == -1) scale_factor = 2;
    const int nw = originalWidth * scale_factor, nh = originalHeight * scale_factor;
    Mat_m u(nh, nw);
    for (int i = 0; i < originalHeight; ++i) for (int j = 0; j < originalWidth; ++j)
      for (int a = 0; a < scale_factor; ++a) for (int b = 0; b < scale_factor; ++b) u(i * scale_factor + a, j * scale_factor + b) = image_matrix(i, j);
    image_matrix = u; originalWidth = nw; originalHeight = nh;
  }
  // image_com.cpp:66-124: five int sizes, then every entry of U, S, V truncated to int and stored as its low byte
  void save_compressed(const std::string& filename) {
    std::ofstream file(filename, std::ios::binary);
    if (!file.is_open()) { std::cerr << "Error opening file: " << filename << std::endl; return; }
    const int dims[5] = {static_cast<int>(left_singular.rows()), static_cast<int>(left_singular.cols()), static_cast<int>(singular.size()),
                         static_cast<int>(right_singular.rows()), static_cast<int>(right_singular.cols())};
    file.write(reinterpret_cast<const char*>(dims), sizeof(dims));
    auto put = [&](double x) { const int v = static_cast<int>(x); const char b = static_cast<char>(v & 0xFF); file.write(&b, 1); };
    for (int i = 0; i < dims[0]; ++i) for (int j = 0; j < dims[1]; ++j) put(left_singular(i, j));
    for (int i = 0; i < dims[2]; ++i) put(singular(i));
    for (int i = 0; i < dims[3]; ++i) for (int j = 0; j < dims[4]; ++j) put(right_singular(i, j));
  }
  // image_com.cpp:131-181
  void load_compressed(const std::string& filename) {
    std::ifstream file(filename, std::ios::binary);
    if (!file.is_open()) { std::cerr << "Error opening file: " << filename << std::endl; return; }
    int dims[5];
    file.read(reinterpret_cast<char*>(dims), sizeof(dims));
    left_singular = Mat_m(dims[0], dims[1]); singular = Vec_v::Zero(dims[2]); right_singular = Mat_m(dims[3], dims[4]);
    auto get = [&]() { char b; file.read(&b, 1); return static_cast<double>(static_cast<unsigned char>(b)); };
    for (int i = 0; i < dims[0]; ++i) for (int j = 0; j < dims[1]; ++j) left_singular(i, j) = get();
    for (int i = 0; i < dims[2]; ++i) singular(i) = get();
    for (int i = 0; i < dims[3]; ++i) for (int j = 0; j < dims[4]; ++j) right_singular(i, j) = get();
  }
  // image_com.cpp:406-411
  double get_compression_ratio() {
    const double initial_size = originalHeight * originalWidth;
    const double compressed_size = degree * (originalWidth + originalHeight + 1);
    return initial_size / compressed_size;
  }

  static uint64_t& seed() { static uint64_t s = 0x5eedULL; return s; }   // additive: the sketch / power start vectors are seeded

 private:
  void compress_impl(int k, int normalize_first) {
    rsvdb_ctx* c = rsvdb::default_context();
    const std::ptrdiff_t m = image_matrix.rows(), n = image_matrix.cols();
    const int kk = k == -1 ? static_cast<int>((m < n ? m : n) / 4) : k;
    const int l = kk + 10;
    Mat_m U(m, l), V(n, l); Vec_v S = Vec_v::Zero(l);
    double lo = 0.0, hi = 0.0;
    rsvdb::check(c, rsvdb_image_compress_host(c, image_matrix.data(), m, n, m, k, normalize_first, nullptr, 0, seed(), &lo, &hi, U.data(), m,
                                              S.data(), V.data(), n, &degree));
    if (normalize_first) { original_min = lo; original_max = hi; }
    left_singular = U; singular = S; right_singular = V;
  }
  Mat_m reconstruct_impl(int denormalize) {
    rsvdb_ctx* c = rsvdb::default_context();
    const std::ptrdiff_t m = left_singular.rows(), n = right_singular.rows();
    Mat_m out(m, n);
    rsvdb::check(c, rsvdb_image_reconstruct_host(c, left_singular.data(), m, m, singular.data(), right_singular.data(), n, n,
                                                 static_cast<int>(singular.size()), denormalize, original_min, original_max, out.data(), m));
    return out;
  }
};

#endif  // IMAGE_COMP_HPP
