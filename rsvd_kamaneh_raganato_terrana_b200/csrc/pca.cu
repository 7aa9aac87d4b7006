// PCA pre/post-passes (reference PCA/include/PCA_class.hpp:24-47,93-100).  All of them are HBM-bound streaming kernels over
// column-major data: a column is contiguous, so one CTA owns one column at a time (grid-stride over columns), reads it
// with coalesced 16-byte loads and reduces in a fixed order -- results do not depend on the launch geometry.
#include "pca.cuh"

#include <algorithm>

#include "comm.cuh"

namespace rsvdb {
namespace {

constexpr int kThreads = 512;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// out[j] = sum_i f(A(i,j)),  f(x) = w_i * x            (mode 0; w == nullptr -> 1)
//                            f(x) = (x - shift[j])^2   (mode 1)
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_col_reduce(const double* __restrict__ A, long long ld, long long rows, int cols,
                                                         const double* __restrict__ aux, double* __restrict__ out) {
  __shared__ double red[33];
  for (int j = blockIdx.x; j < cols; j += gridDim.x) {
    const double* col = A + (size_t)j * ld;
    const double sh = (MODE == 1) ? aux[j] : 0.0;
    double acc0 = 0.0, acc1 = 0.0;
    // columns start 16-byte aligned when ld is even and the base is; otherwise peel one element
    long long i0 = ((reinterpret_cast<uintptr_t>(col) & 15) != 0) ? 1 : 0;
    if (i0 == 1 && threadIdx.x == 0 && rows > 0) {
      const double x = col[0];
      acc0 = (MODE == 1) ? (x - sh) * (x - sh) : (aux ? aux[0] * x : x);
    }
    const long long pairs = rows > i0 ? (rows - i0) >> 1 : 0;
    const double2* c2 = reinterpret_cast<const double2*>(col + i0);
    for (long long p = threadIdx.x; p < pairs; p += kThreads) {
      const double2 v = c2[p];
      if (MODE == 1) { acc0 += (v.x - sh) * (v.x - sh); acc1 += (v.y - sh) * (v.y - sh); }
      else if (aux) { acc0 += aux[i0 + 2 * p] * v.x; acc1 += aux[i0 + 2 * p + 1] * v.y; }
      else { acc0 += v.x; acc1 += v.y; }
    }
    const long long tail = i0 + 2 * pairs;
    if (tail < rows && threadIdx.x == 1) {
      const double x = col[tail];
      acc1 += (MODE == 1) ? (x - sh) * (x - sh) : (aux ? aux[tail] * x : x);
    }
    const double s = block_sum(acc0 + acc1, red);
    if (threadIdx.x == 0) out[j] = s;
  }
}

// stats[0..n) holds sums, stats[n] the global row count; -> mean.  Second call: css -> stddev, inv_sd.
__global__ void k_finish_mean(const double* __restrict__ sums, int n, double* __restrict__ mean) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) mean[j] = sums[j] / sums[n];
}
__global__ void k_finish_sd(const double* __restrict__ css, double rows_total, int n, double* __restrict__ sd, double* __restrict__ inv_sd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double s = sqrt(css[j] / (rows_total - 1.0));
  if (sd) sd[j] = s;
  if (inv_sd) inv_sd[j] = 1.0 / s;
}
__global__ void k_set(double* p, double v) { *p = v; }

// A(i,j) = (A(i,j) - mean[j]) / sd[j]      (a true division, like the reference's `array().rowwise() /= stddev_`)
__global__ void __launch_bounds__(kThreads) k_center_scale(double* __restrict__ A, long long ld, long long rows, int cols,
                                                           const double* __restrict__ mean, const double* __restrict__ sd) {
  for (int j = blockIdx.x; j < cols; j += gridDim.x) {
    double* col = A + (size_t)j * ld;
    const double mu = mean[j], dv = sd ? sd[j] : 1.0;
    long long i0 = ((reinterpret_cast<uintptr_t>(col) & 15) != 0) ? 1 : 0;
    if (i0 == 1 && threadIdx.x == 0 && blockIdx.y == 0 && rows > 0) col[0] = sd ? (col[0] - mu) / dv : col[0] - mu;
    const long long pairs = rows > i0 ? (rows - i0) >> 1 : 0;
    double2* c2 = reinterpret_cast<double2*>(col + i0);
    for (long long p = (long long)blockIdx.y * kThreads + threadIdx.x; p < pairs; p += (long long)gridDim.y * kThreads) {
      double2 v = c2[p];
      if (sd) { v.x = (v.x - mu) / dv; v.y = (v.y - mu) / dv; } else { v.x -= mu; v.y -= mu; }
      c2[p] = v;
    }
    const long long tail = i0 + 2 * pairs;
    if (tail < rows && threadIdx.x == 1 && blockIdx.y == 0) col[tail] = sd ? (col[tail] - mu) / dv : col[tail] - mu;
  }
}

// element-wise updates on skinny (rows x cols) operands; thread = row, loop over a column slab
__global__ void k_scale_rows_copy(const double* __restrict__ X, long long ldx, double* __restrict__ Xs, long long lds, long long n,
                                  int l, const double* __restrict__ inv_sd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = inv_sd ? inv_sd[i] : 1.0;
  for (int j = blockIdx.y; j < l; j += gridDim.y) Xs[(size_t)j * lds + i] = X[(size_t)j * ldx + i] * s;
}
__global__ void k_sub_col_const(double* __restrict__ Y, long long ld, long long rows, int cols, const double* __restrict__ w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) Y[(size_t)j * ld + i] -= w[j];
}
__global__ void k_rank1_correct(double* __restrict__ Z, long long ld, long long rows, int cols, const double* __restrict__ mu,
                                const double* __restrict__ s, const double* __restrict__ inv_sd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const double m = mu[i], d = inv_sd ? inv_sd[i] : 1.0;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) {
    const double z = Z[(size_t)j * ld + i] - m * s[j];
    Z[(size_t)j * ld + i] = inv_sd ? z * d : z;
  }
}
__global__ void k_add_row_vector(double* __restrict__ out, long long ld, long long rows, long long cols, const double* __restrict__ mean,
                                 double sign) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (long long j = blockIdx.y; j < cols; j += gridDim.y) out[(size_t)j * ld + i] += sign * mean[j];
}

inline dim3 skinny_grid(int64_t rows, int64_t cols) {
  return dim3((unsigned)((rows + 255) / 256), (unsigned)std::max<int64_t>(1, std::min<int64_t>(cols, 32)));
}

}  // namespace

int weighted_colsum(rsvdb_ctx* c, const double* X, int64_t ld, int64_t rows, int cols, const double* wgt, double* out) {
  if (cols <= 0) return 0;
  k_col_reduce<0><<<std::min(cols, 8 * c->nsm), kThreads, 0, c->stream>>>(X, ld, rows, cols, wgt, out);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}

int column_stats(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, double* mean, double* stddev, double* inv_sd) {
  if (n <= 0 || n > INT32_MAX) return fail(c, -1, "column_stats: bad column count");
  PhaseTimer pt(c, PH_OTHER);
  RSVDB_CUDA(c, c->pca_ws.reserve(std::max<size_t>(c->pca_ws.bytes, (size_t)(2 * n + 8) * sizeof(double))));
  double* sums = c->pca_ws.ptr;            // [n sums | row count]
  double* css = sums + n + 1;              // [n centred sums of squares]
  const int grid = (int)std::min<int64_t>(n, 8 * c->nsm);
  k_col_reduce<0><<<grid, kThreads, 0, c->stream>>>(A, lda, m, (int)n, nullptr, sums);
  k_set<<<1, 1, 0, c->stream>>>(sums + n, (double)m);
  RSVDB_CUDA(c, cudaGetLastError()); c->launches += 2;
  if (c->nranks > 1) RSVDB_TRY(comm_allreduce_sum(c, sums, (size_t)n + 1));
  k_finish_mean<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(sums, (int)n, mean);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  if (stddev || inv_sd) {
    // second sweep over the column (the reference squares the CENTRED data, PCA_class.hpp:39); mostly L2 hits when the
    // shard fits the 126 MB L2, one more HBM pass otherwise
    k_col_reduce<1><<<grid, kThreads, 0, c->stream>>>(A, lda, m, (int)n, mean, css);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
    double rows_total = (double)m;
    if (c->nranks > 1) {
      RSVDB_TRY(comm_allreduce_sum(c, css, (size_t)n));
      RSVDB_CUDA(c, cudaMemcpyAsync(&rows_total, sums + n, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    k_finish_sd<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(css, rows_total, (int)n, stddev, inv_sd);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  }
  return 0;
}

int center_columns(rsvdb_ctx* c, double* A, int64_t m, int64_t n, int64_t lda, const double* mean, const double* sd) {
  if (n <= 0 || m <= 0) return 0;
  PhaseTimer pt(c, PH_OTHER);
  // enough CTAs to saturate HBM even for a handful of long columns: split the rows over gridDim.y
  const int gx = (int)std::min<int64_t>(n, 4 * c->nsm);
  const int gy = (int)std::max<int64_t>(1, std::min<int64_t>((4 * c->nsm + gx - 1) / gx, (m / 2 + kThreads - 1) / kThreads));
  k_center_scale<<<dim3(gx, gy), kThreads, 0, c->stream>>>(A, lda, m, (int)n, mean, sd);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}

int scale_rows_copy(rsvdb_ctx* c, const double* X, int64_t ldx, double* Xs, int64_t lds, int64_t n, int l, const double* inv_sd) {
  if (n <= 0 || l <= 0) return 0;
  k_scale_rows_copy<<<skinny_grid(n, l), 256, 0, c->stream>>>(X, ldx, Xs, lds, n, l, inv_sd);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}
int sub_col_const(rsvdb_ctx* c, double* Y, int64_t ld, int64_t rows, int cols, const double* w) {
  if (rows <= 0 || cols <= 0) return 0;
  k_sub_col_const<<<skinny_grid(rows, cols), 256, 0, c->stream>>>(Y, ld, rows, cols, w);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}
int rank1_correct(rsvdb_ctx* c, double* Z, int64_t ld, int64_t rows, int cols, const double* mu, const double* s, const double* inv_sd) {
  if (rows <= 0 || cols <= 0) return 0;
  k_rank1_correct<<<skinny_grid(rows, cols), 256, 0, c->stream>>>(Z, ld, rows, cols, mu, s, inv_sd);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}
int add_row_vector(rsvdb_ctx* c, double* out, int64_t ld, int64_t rows, int64_t cols, const double* mean, double sign) {
  if (rows <= 0 || cols <= 0) return 0;
  k_add_row_vector<<<skinny_grid(rows, cols), 256, 0, c->stream>>>(out, ld, rows, cols, mean, sign);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  return 0;
}

}  // namespace rsvdb
