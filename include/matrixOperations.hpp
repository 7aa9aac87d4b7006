// Drop-in for reference include/matrixOperations.hpp: manualMatrixMultiply (src/matrixOperations.cpp:7-28).
#ifndef matrixOperations_H
#define matrixOperations_H

#include "rsvdb_dense.hpp"

using Mat = Mat_m;
using Vec = Vec_v;

// Throws std::invalid_argument on a shape mismatch like the reference (:8-11); the product runs on the GPU.
inline Mat manualMatrixMultiply(const Mat& matrix1, const Mat& matrix2) {
  rsvdb_ctx* c = rsvdb::default_context();
  Mat result(matrix1.rows(), matrix2.cols());
  rsvdb::check(c, rsvdb_gemm_host(c, matrix1.data(), matrix1.rows(), matrix1.cols(), matrix1.rows() > 0 ? matrix1.rows() : 1, matrix2.data(),
                                  matrix2.rows(), matrix2.cols(), matrix2.rows() > 0 ? matrix2.rows() : 1, result.data(),
                                  matrix1.rows() > 0 ? matrix1.rows() : 1));
  return result;
}

#endif
