// K8: SVDMethod::Power -- dominant singular triplets by power iteration with rank-1 deflation.
//
// Replaces SVD<Power>::powerMethodSVD (reference include/SVD_class.hpp:184-219) and PM (src/PM.cpp:4-81).  The reference
// forms the n x n Gram matrix B = M^T M (:193), runs s = ceil(log(4 log(2n/delta)/(eps delta))/(2 lambda)) iterations
// x <- B x / ||B x|| with an MPI Gatherv + Bcast per iteration (PM.cpp:25-28,40-69), then v = x, sigma = ||M v||,
// u = M v / sigma (:72-79) and deflates M -= sigma u v^T, B -= (sigma u v^T)^T (sigma u v^T) (:210-212).
// Here B is never formed: x <- M^T (M x) on the deflated M is the same iteration (B x = M^T (M x)), costs 2 passes over
// the r x c factor instead of one over c x c, and keeps everything in L2.  M is held TRANSPOSED (Mt, c x r column-major),
// which is how the pipeline produces it (B^T = A^T Q), so both products read memory contiguously.
// The start vector is drawn from a counter-based generator seeded by the caller (the reference: std::random_device).
#include "pipeline.cuh"

#include <cmath>

namespace rsvdb {

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double std_normal(uint64_t seed, uint64_t idx) {
  const uint64_t a = splitmix64(seed ^ splitmix64(2 * idx)), b = splitmix64(seed ^ splitmix64(2 * idx + 1));
  const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  __syncthreads();
  return t;
}

// x_j = N(0,1); partial[b] = sum of squares of block b
__global__ void k_pm_start(double* x, int64_t cdim, uint64_t seed, uint64_t stream_id, double* partial, const int* stop) {
  __shared__ double red[32];
  if (*stop) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (j < cdim) { v = std_normal(seed + 0x632BE59BD9B4E019ull * stream_id, (uint64_t)j); x[j] = v; }
  const double s = block_sum(v * v, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// y_i = dot(Mt[:, i], x) / ||x||,  ||x||^2 = sum(partial[0..np))          (one CTA per row i of M)
__global__ void k_pm_mx(const double* __restrict__ Mt, int64_t ldmt, int64_t cdim, const double* __restrict__ x,
                        const double* __restrict__ partial, int np, double* __restrict__ y, const int* stop) {
  __shared__ double red[32];
  if (*stop) return;
  double nrm2 = 0.0;
  for (int b = 0; b < np; ++b) nrm2 += partial[b];
  const double inv = (np < 0) ? 1.0 : (nrm2 > 0.0 ? 1.0 / sqrt(nrm2) : 0.0);   // np < 0: x is already a unit vector
  const double* col = Mt + (size_t)blockIdx.x * ldmt;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < cdim; j += blockDim.x) acc = fma(col[j], x[j], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) y[blockIdx.x] = acc * inv;
}
// z_j = sum_i Mt[j][i] y_i ; partial[b] = block sum of z_j^2
__global__ void k_pm_mty(const double* __restrict__ Mt, int64_t ldmt, int64_t cdim, int r, const double* __restrict__ y,
                         double* __restrict__ z, double* __restrict__ partial, const int* stop) {
  __shared__ double red[32];
  constexpr int YT = 2048;                       // y is staged through shared memory in fixed tiles: any r launches
  __shared__ double ys[YT];
  if (*stop) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  for (int i0 = 0; i0 < r; i0 += YT) {
    const int nt = min(YT, r - i0);
    __syncthreads();
    for (int i = threadIdx.x; i < nt; i += blockDim.x) ys[i] = y[i0 + i];
    __syncthreads();
    if (j < cdim) for (int i = 0; i < nt; ++i) acc = fma(Mt[(size_t)(i0 + i) * ldmt + j], ys[i], acc);
  }
  if (j < cdim) z[j] = acc;
  const double s = block_sum(acc * acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// v = x / ||x|| -> V[:, idx]                                                                       (PM.cpp:72-73)
__global__ void k_pm_store_v(const double* __restrict__ x, const double* __restrict__ partial, int np, int64_t cdim,
                             double* __restrict__ vout, const int* stop) {
  if (*stop) return;
  double nrm2 = 0.0;
  for (int b = 0; b < np; ++b) nrm2 += partial[b];
  const double inv = nrm2 > 0.0 ? 1.0 / sqrt(nrm2) : 0.0;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cdim) vout[j] = x[j] * inv;
}
// sigma = ||y||, u = y / sigma (PM.cpp:76-79); early exit when sigma < 1e-12 (SVD_class.hpp:198)
__global__ void k_pm_sigma(const double* __restrict__ y, int r, double* __restrict__ ucol, double* __restrict__ S, int idx,
                           int* stop, int* found, double* sigma_out) {
  __shared__ double red[32];
  if (*stop) return;
  double acc = 0.0;
  for (int i = threadIdx.x; i < r; i += blockDim.x) acc = fma(y[i], y[i], acc);
  const double sg = sqrt(block_sum(acc, red));
  if (sg < 1e-12) { if (threadIdx.x == 0) *stop = 1; return; }
  for (int i = threadIdx.x; i < r; i += blockDim.x) ucol[i] = y[i] / sg;
  if (threadIdx.x == 0) { S[idx] = sg; *found = idx + 1; *sigma_out = sg; }
}
// Mt[j][i] -= sigma * v_j * u_i                                                               (SVD_class.hpp:210-211)
__global__ void k_pm_deflate(double* __restrict__ Mt, int64_t ldmt, int64_t cdim, int r, const double* __restrict__ v,
                             const double* __restrict__ ucol, const double* __restrict__ sigma, const int* stop) {
  if (*stop) return;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cdim) return;
  const double sv = (*sigma) * v[j];
  for (int i = blockIdx.y; i < r; i += gridDim.y) Mt[(size_t)i * ldmt + j] -= sv * ucol[i];
}
__global__ void k_set_identity(double* __restrict__ U, int64_t ldu, int r, int cols) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < r * cols) { const int i = e % r, k = e / r; U[(size_t)k * ldu + i] = (i == k) ? 1.0 : 0.0; }
}

}  // namespace

int pm_iterations(int64_t ncols) {   // src/PM.cpp:25-28
  const double epsilon = 1.e-10, delta = 0.05, lambda = 0.1;
  return (int)std::ceil(std::log(4 * std::log(2 * ncols / delta) / (epsilon * delta)) / (2 * lambda));
}

// Mt: c x r (column-major, ldmt), the TRANSPOSE of the r x c data matrix; it is deflated in place.
// U: r x r, identity-initialised like SVD::compute() (:82), column i <- u_i.  S: min(r,c), zero-initialised.
// U has u_cols identity-initialised columns (r for SVD<Power>, 1 for a bare PM call).
// V: c x dim, column i <- v_i (the drop-in wrappers re-shape this to the reference's rows-of-an-n-x-n-identity layout).
int small_svd_power_t(rsvdb_ctx* c, double* Mt, int64_t ldmt, int64_t r, int64_t cdim, int rdim, uint64_t seed,
                      double* U, int64_t ldu, int u_cols, double* S, double* V, int64_t ldv, int* found_host) {
  PhaseTimer pt(c, PH_SMALL_SVD);
  c->d_svd_info = nullptr;                  // no Jacobi sweep count belongs to this call
  const int64_t kmin = std::min(r, cdim);
  const int dim = rdim ? rdim : (int)kmin;
  if (dim > kmin) return fail(c, -1, "SVD<Power>: r larger than min(rows, cols)");
  const int threads = 256;
  const int nb = (int)((cdim + threads - 1) / threads);
  const size_t need = ((size_t)2 * cdim + r + nb + 8) * sizeof(double) + 64;
  RSVDB_CUDA(c, c->qr2_ws.reserve(need));   // qr2_ws is idle during the small SVD (svd_ws may hold the caller's operands)
  double* x = c->qr2_ws.ptr; double* z = x + cdim; double* y = z + cdim; double* partial = y + r;
  double* sigma = partial + nb; int* flags = reinterpret_cast<int*>(sigma + 2);   // flags[0] = stop, flags[1] = found
  cudaStream_t st = c->stream;
  RSVDB_CUDA(c, cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));
  RSVDB_CUDA(c, cudaMemsetAsync(S, 0, (size_t)kmin * sizeof(double), st));
  k_set_identity<<<(int)((r * u_cols + 255) / 256), 256, 0, st>>>(U, ldu, (int)r, u_cols);
  const int s = pm_iterations(cdim);
  int64_t nl = 1;
  for (int i = 0; i < dim; ++i) {
    k_pm_start<<<nb, threads, 0, st>>>(x, cdim, seed, (uint64_t)i, partial, flags);
    double* cur = x; double* nxt = z;
    for (int it = 0; it < s; ++it) {
      k_pm_mx<<<(int)r, threads, 0, st>>>(Mt, ldmt, cdim, cur, partial, nb, y, flags);
      k_pm_mty<<<nb, threads, 0, st>>>(Mt, ldmt, cdim, (int)r, y, nxt, partial, flags);
      std::swap(cur, nxt);
    }
    double* vcol = V + (size_t)i * ldv;
    k_pm_store_v<<<nb, threads, 0, st>>>(cur, partial, nb, cdim, vcol, flags);
    k_pm_mx<<<(int)r, threads, 0, st>>>(Mt, ldmt, cdim, vcol, partial, -1, y, flags);    // y = M v
    nl += 2 * s + 3;
    k_pm_sigma<<<1, 256, 0, st>>>(y, (int)r, U + (size_t)i * ldu, S, i, flags, flags + 1, sigma);
    dim3 g(nb, (unsigned)std::min<int64_t>(r, 64));
    k_pm_deflate<<<g, threads, 0, st>>>(Mt, ldmt, cdim, (int)r, vcol, U + (size_t)i * ldu, sigma, flags);
    nl += 2;
  }
  RSVDB_CUDA(c, cudaGetLastError());
  c->launches += nl;
  if (found_host) {
    RSVDB_CUDA(c, cudaStreamSynchronize(st));
    RSVDB_CUDA(c, cudaMemcpy(found_host, flags + 1, sizeof(int), cudaMemcpyDeviceToHost));
  }
  return 0;
}

}  // namespace rsvdb
