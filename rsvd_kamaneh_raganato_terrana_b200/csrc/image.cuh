// Image::normalize / compress / reconstruct / deNormalize around the rSVD path (reference image_compression/src/image_com.cpp).
#pragma once
#include <cstdint>
#include "context.cuh"
#include "pipeline.cuh"

namespace rsvdb {
// d_mm[0] = min, d_mm[1] = max over the m x n matrix (image_com.cpp:253-254); two-stage, fixed order
int image_minmax(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, double* d_mm);
// in place: (x - min) / (max - min) (:257-260) or its inverse x * (max - min) + min (:272-275); a no-op when min >= max
int image_affine(rsvdb_ctx* c, double* A, int64_t m, int64_t n, int64_t lda, const double* d_mm, bool inverse);
// out (m x n) = U diag(S) V^T (:184-190), followed by the inverse affine map when d_mm != nullptr
int image_reconstruct(rsvdb_ctx* c, const double* U, int64_t m, int64_t ldu, const double* S, const double* V, int64_t n, int64_t ldv, int l,
                      const double* d_mm, double* out, int64_t ldout);
}  // namespace rsvdb
