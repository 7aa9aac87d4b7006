// Image::normalize / compress / reconstruct / deNormalize (reference image_compression/src/image_com.cpp:184-190,251-317)
// as one device pipeline around the rSVD path: a min-max reduction and an affine map before (HBM-bound, one read + one
// read-modify-write of the image), the older-API rSVD (q = 1, Power back-end, l = k + 10) in the middle, and
// U diag(S) V^T with the inverse affine map fused into a single pass over the reconstructed image.
#include "image.cuh"

#include <algorithm>
#include <cfloat>

namespace rsvdb {
namespace {

constexpr int kT = 512;

__device__ __forceinline__ void warp_minmax(double& lo, double& hi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
}
// stage 1: one partial (min, max) per CTA over a column-major m x n matrix with leading dimension ld
__global__ void __launch_bounds__(kT) k_minmax_partial(const double* __restrict__ A, long long ld, long long rows, long long cols,
                                                       double* __restrict__ part) {
  __shared__ double slo[kT / 32], shi[kT / 32];
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (long long j = blockIdx.x; j < cols; j += gridDim.x) {
    const double* col = A + (size_t)j * ld;
    for (long long i = threadIdx.x; i < rows; i += kT) { const double x = col[i]; lo = fmin(lo, x); hi = fmax(hi, x); }
  }
  warp_minmax(lo, hi);
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = threadIdx.x < kT / 32 ? slo[threadIdx.x] : DBL_MAX; hi = threadIdx.x < kT / 32 ? shi[threadIdx.x] : -DBL_MAX;
    warp_minmax(lo, hi);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = lo; part[2 * blockIdx.x + 1] = hi; }
  }
}
// stage 2: mm[0] = min, mm[1] = max
__global__ void k_minmax_final(const double* __restrict__ part, int n, double* __restrict__ mm) {
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (int i = threadIdx.x; i < n; i += 32) { lo = fmin(lo, part[2 * i]); hi = fmax(hi, part[2 * i + 1]); }
  warp_minmax(lo, hi);
  if (threadIdx.x == 0) { mm[0] = lo; mm[1] = hi; }
}
// forward: x <- (x - min) / (max - min) when min < max (image_com.cpp:257-260); inverse: x <- x * (max - min) + min (:272-275)
__global__ void __launch_bounds__(kT) k_affine(double* __restrict__ A, long long ld, long long rows, long long cols,
                                               const double* __restrict__ mm, int inverse) {
  const double lo = mm[0], hi = mm[1];
  if (!(lo < hi)) return;
  const double range = hi - lo;
  for (long long j = blockIdx.x; j < cols; j += gridDim.x) {
    double* col = A + (size_t)j * ld;
    for (long long i = (long long)blockIdx.y * kT + threadIdx.x; i < rows; i += (long long)gridDim.y * kT)
      col[i] = inverse ? col[i] * range + lo : (col[i] - lo) / range;
  }
}
// Us(:, j) = U(:, j) * S[j]
__global__ void k_scale_cols(const double* __restrict__ U, long long ldu, double* __restrict__ Us, long long lds, long long rows, int cols,
                             const double* __restrict__ S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) Us[(size_t)j * lds + i] = U[(size_t)j * ldu + i] * S[j];
}

inline dim3 col_grid(rsvdb_ctx* c, int64_t rows, int64_t cols) {
  const int gx = (int)std::min<int64_t>(cols, 4 * c->nsm);
  const int gy = (int)std::max<int64_t>(1, std::min<int64_t>((4 * c->nsm + gx - 1) / gx, (rows + kT - 1) / kT));
  return dim3(gx, gy);
}

}  // namespace

int image_minmax(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, double* d_mm) {
  PhaseTimer pt(c, PH_OTHER);
  const int grid = (int)std::min<int64_t>(n, 4 * c->nsm);
  RSVDB_CUDA(c, c->pca_ws.reserve(std::max<size_t>(c->pca_ws.bytes, (size_t)(2 * grid + 8) * sizeof(double))));
  double* part = c->pca_ws.ptr;
  k_minmax_partial<<<grid, kT, 0, c->stream>>>(A, lda, m, n, part);
  k_minmax_final<<<1, 32, 0, c->stream>>>(part, grid, d_mm);
  RSVDB_CUDA(c, cudaGetLastError()); c->launches += 2;
  return 0;
}

int image_affine(rsvdb_ctx* c, double* A, int64_t m, int64_t n, int64_t lda, const double* d_mm, bool inverse) {
  PhaseTimer pt(c, PH_OTHER);
  k_affine<<<col_grid(c, m, n), kT, 0, c->stream>>>(A, lda, m, n, d_mm, inverse ? 1 : 0);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}

int image_reconstruct(rsvdb_ctx* c, const double* U, int64_t m, int64_t ldu, const double* S, const double* V, int64_t n, int64_t ldv, int l,
                      const double* d_mm, double* out, int64_t ldout) {
  // scratch in tmp_ws: Us (m x l), Vt (l x n)
  const int64_t ldl = std::max<int64_t>(2, (l + 1) & ~1);
  RSVDB_CUDA(c, c->tmp_ws.reserve(std::max<size_t>(c->tmp_ws.bytes, ((size_t)m * l + (size_t)ldl * n + 64) * sizeof(double))));
  double* Us = c->tmp_ws.ptr; double* Vt = Us + (((size_t)m * l + 31) & ~size_t(31));
  {
    PhaseTimer pt(c, PH_OTHER);
    k_scale_cols<<<dim3((unsigned)((m + 255) / 256), (unsigned)std::min(l, 32)), 256, 0, c->stream>>>(U, ldu, Us, m, m, l, S);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  }
  RSVDB_TRY(transpose2d(c, V, ldv, Vt, ldl, n, l));
  {
    PhaseTimer pt(c, PH_GEMM_AN);
    int nl = 0;
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, Us, m, l, m, Vt, ldl, (int)n, out, ldout, &nl));
    c->launches += nl;
  }
  if (d_mm) RSVDB_TRY(image_affine(c, out, m, n, ldout, d_mm, true));
  return 0;
}

}  // namespace rsvdb
