// C ABI (include/rsvdb.h): context management, device entry points and the host-pointer mirrors of the reference API.
#include "../../include/rsvdb.h"

#include <algorithm>
#include <vector>

#include "comm.cuh"
#include "context.cuh"
#include "image.cuh"
#include "pca.cuh"
#include "pipeline.cuh"
#include "pod.cuh"
#include "spmm.cuh"
#include "tsqr.cuh"

using namespace rsvdb;

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// K9: Omega(i, j) ~ N(0,1), a pure function of (seed, i, j): identical on every rank and for every launch geometry.
__global__ void k_fill_normal(double* __restrict__ out, long long ld, long long rows, int cols, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) {
    const uint64_t idx = (uint64_t)j * (uint64_t)rows + (uint64_t)i;
    const uint64_t a = mix64(seed ^ mix64(2 * idx)), b = mix64(seed ^ mix64(2 * idx + 1));
    const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
    out[(size_t)j * ld + i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
}
// Givens sign convention for the QR API: make diag(R) >= 0 by flipping row i of R and column i of Q together.
__global__ void k_sign_normalise(double* __restrict__ Q, long long ldq, long long qrows, double* __restrict__ R, long long ldr,
                                 int rcols, int k) {
  const int i = blockIdx.x;   // reflector / diagonal index
  if (i >= k) return;
  const double d = R[(size_t)i * ldr + i];
  if (!(d < 0.0)) return;
  for (int j = threadIdx.x; j < rcols; j += blockDim.x) R[(size_t)j * ldr + i] = -R[(size_t)j * ldr + i];
  for (long long r = threadIdx.x; r < qrows; r += blockDim.x) Q[(size_t)i * ldq + r] = -Q[(size_t)i * ldq + r];
}

__global__ void k_reciprocal(const double* __restrict__ x, double* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = 1.0 / x[i];
}


int h2d(rsvdb_ctx* c, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  PhaseTimer pt(c, PH_COPY);
  RSVDB_CUDA(c, upload_block(c, c->stream, dst, ldd, src, lds, rows, cols));
  return 0;
}
int d2h(rsvdb_ctx* c, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  PhaseTimer pt(c, PH_COPY);
  if (ldd == rows && lds == rows) {
    RSVDB_CUDA(c, cudaMemcpyAsync(dst, src, (size_t)rows * cols * 8, cudaMemcpyDeviceToHost, c->stream));
  } else {
    RSVDB_CUDA(c, cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyDeviceToHost, c->stream));
  }
  return 0;
}

// Host entry points synchronise anyway: report a Jacobi SVD that hit the sweep cap (negative sweep count) instead of
// returning RSVDB_OK with unconverged factors.  (_dev callers query rsvdb_last_svd_info themselves.)
int check_svd_converged(rsvdb_ctx* c) {
  if (!c->d_svd_info) return 0;
  int h[2] = {0, 0};
  RSVDB_CUDA(c, cudaMemcpy(h, c->d_svd_info, sizeof(h), cudaMemcpyDeviceToHost));
  if (h[0] < 0) return fail(c, RSVDB_ERR_NO_CONVERGENCE, "Jacobi SVD: sweep cap reached without convergence");
  return 0;
}

// bump allocator over io_ws
struct IoArena {
  rsvdb_ctx* c; size_t off = 0;
  explicit IoArena(rsvdb_ctx* ctx) : c(ctx) {}
  static size_t pad(size_t doubles) { return (doubles + 31) & ~size_t(31); }
  double* take(size_t doubles) { double* p = c->io_ws.ptr + off; off += pad(doubles); return p; }
};

}  // namespace

extern "C" {

const char* rsvdb_version(void) { return "rsvdb 0.1 (sm_100a)"; }

int rsvdb_create(rsvdb_ctx** out, int device) {
  if (!out) return RSVDB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return RSVDB_ERR_CUDA;
  rsvdb_ctx* c = new rsvdb_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  if (prop.major != 10) { delete c; return RSVDB_ERR_UNSUPPORTED; }   // sm_100a only: no other code path exists
  c->nsm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  c->stream = c->own_stream;
  if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  for (auto& e : c->side_ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  *out = c;
  return RSVDB_OK;
}

int rsvdb_destroy(rsvdb_ctx* c) {
  if (!c) return RSVDB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
  for (auto e : c->side_ev) if (e) cudaEventDestroy(e);
  for (auto& s : c->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  c->gemm_ws.release(); c->qr_ws.release(); c->qr2_ws.release(); c->tmp_ws.release(); c->svd_ws.release(); c->io_ws.release(); c->wide_ws.release(); c->pca_ws.release(); c->pod_ws.release(); c->chol_ws.release();
  if (c->chol_host) cudaFreeHost(c->chol_host);
  delete c->stager;
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return RSVDB_OK;
}

int rsvdb_set_stream(rsvdb_ctx* c, void* s) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  c->stream = static_cast<cudaStream_t>(s);
  return RSVDB_OK;
}
int rsvdb_use_own_stream(rsvdb_ctx* c) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  c->stream = c->own_stream;
  return RSVDB_OK;
}
int rsvdb_synchronize(rsvdb_ctx* c) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}
const char* rsvdb_last_error(const rsvdb_ctx* c) { return c ? c->err.c_str() : "null context"; }
int64_t rsvdb_launch_count(const rsvdb_ctx* c) { return c ? c->launches : 0; }
int64_t rsvdb_generic_gemm_fallbacks(void) { return generic_fallback_count(); }
int64_t rsvdb_split_gemm_products(void) { return split_product_count(); }
int rsvdb_set_qr_policy(rsvdb_ctx* c, int policy) {
  if (!c || (policy != 0 && policy != 1)) return RSVDB_ERR_INVALID_ARGUMENT;
  c->qr_policy = policy;
  return RSVDB_OK;
}
int rsvdb_qr_path_counts(const rsvdb_ctx* c, int64_t* cholqr2, int64_t* householder) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (cholqr2) *cholqr2 = c->qr_fast;
  if (householder) *householder = c->qr_householder;
  return RSVDB_OK;
}

int rsvdb_set_profiling(rsvdb_ctx* c, int enabled) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  c->profiling = enabled != 0;
  return RSVDB_OK;
}
int rsvdb_phase_ms(rsvdb_ctx* c, double* out) {
  if (!c || !out) return RSVDB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < RSVDB_NUM_PHASES; ++i) out[i] = 0.0;
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (auto& s : c->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess && s.phase >= 0 && s.phase < RSVDB_NUM_PHASES) out[s.phase] += ms;
    c->event_pool.push_back(s.a); c->event_pool.push_back(s.b);
  }
  c->spans.clear();
  return RSVDB_OK;
}
int rsvdb_last_svd_info(rsvdb_ctx* c, int* sweeps, int* rotations) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  int h[2] = {0, 0};
  if (c->d_svd_info) {
    RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
    RSVDB_CUDA(c, cudaMemcpy(h, c->d_svd_info, sizeof(h), cudaMemcpyDeviceToHost));
  }
  if (sweeps) *sweeps = h[0];
  if (rotations) *rotations = h[1];
  return RSVDB_OK;
}

int rsvdb_comm_unique_id(void* out) {
  if (!out) return RSVDB_ERR_INVALID_ARGUMENT;
  std::string err;
  return comm_unique_id(out, &err);
}
int rsvdb_comm_init(rsvdb_ctx* c, int nranks, int rank, const void* id) {
  if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "comm_init: bad rank/size");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return comm_init(c, nranks, rank, id);
}
int rsvdb_comm_size(const rsvdb_ctx* c) { return c ? c->nranks : 0; }
int rsvdb_comm_rank(const rsvdb_ctx* c) { return c ? c->rank : 0; }

int rsvdb_pm_iterations(int64_t ncols) { return pm_iterations(ncols); }

// ---------------------------------------------------------------------------------------------------------------------
// device entry points
// ---------------------------------------------------------------------------------------------------------------------
int rsvdb_gemm_an_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dX, int64_t ldx,
                      int l, double* dY, int64_t ldy) {
  if (!c || m < 0 || n < 0 || l < 0 || lda < m || ldx < n || ldy < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "gemm_an: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  int k = 0;
  RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, dA, m, n, lda, dX, ldx, l, dY, ldy, &k));
  c->launches += k;
  return RSVDB_OK;
}

int rsvdb_gemm_at_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dQ, int64_t ldq,
                      int l, double* dZ, int64_t ldz, int transpose_out) {
  if (!c || m < 0 || n < 0 || l < 0 || lda < m || ldq < m || ldz < (transpose_out ? l : n))
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "gemm_at: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  int k = 0;
  RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, dA, m, n, lda, dQ, ldq, l, dZ, ldz, transpose_out, &k));
  c->launches += k;
  return RSVDB_OK;
}

int rsvdb_qr_dev(rsvdb_ctx* c, double* dY, int64_t rows, int l, int64_t ldy, int sharded, double* dR) {
  if (!c || rows < 0 || l <= 0 || ldy < rows) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "qr: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const double* R = nullptr;
  RSVDB_TRY(qr_inplace(c, dY, rows, l, ldy, sharded != 0, &R));
  if (dR) RSVDB_CUDA(c, cudaMemcpyAsync(dR, R, (size_t)l * l * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return RSVDB_OK;
}

int rsvdb_orthonormalize_dev(rsvdb_ctx* c, double* dY, int64_t rows, int l, int64_t ldy, int sharded, double* dR, int* path) {
  if (!c || rows < 0 || l <= 0 || ldy < rows) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "orthonormalize: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const double* R = nullptr;
  const int64_t fast0 = c->qr_fast;
  c->chol_failed = false;
  RSVDB_TRY(orthonormalize(c, dY, rows, l, ldy, sharded != 0, dR ? &R : nullptr));   // without R the second round may take its first-order form
  if (dR) RSVDB_CUDA(c, cudaMemcpyAsync(dR, R, (size_t)l * l * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (path) *path = (c->qr_fast > fast0) ? 0 : 1;
  return RSVDB_OK;
}

int rsvdb_range_finder_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dOmega, int64_t ldo,
                           int l, int q, double* dQ, int64_t ldq) {
  if (!c || m < 0 || n <= 0 || lda < m || ldo < n || ldq < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "range_finder: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return range_finder(c, dA, m, n, lda, dOmega, ldo, l, q, dQ, ldq);
}

int rsvdb_rsvd_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dOmega, int64_t ldo, int l,
                   int q, int method, uint64_t seed, double* dU, int64_t ldu, double* dS, double* dV, int64_t ldv) {
  if (!c || m < 0 || n <= 0 || lda < m || ldo < n || ldu < m || ldv < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rsvd: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return rsvd_device(c, dA, m, n, lda, dOmega, ldo, l, q, method, dU, ldu, dS, dV, ldv, seed);
}

int rsvdb_generate_omega_dev(rsvdb_ctx* c, int64_t n, int l, uint64_t seed, double* dOmega, int64_t ldo) {
  if (!c || n < 0 || l < 0 || ldo < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "generate_omega: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  if (n == 0 || l == 0) return RSVDB_OK;
  dim3 g((unsigned)((n + 255) / 256), (unsigned)std::min(l, 128));
  k_fill_normal<<<g, 256, 0, c->stream>>>(dOmega, ldo, n, l, seed);
  RSVDB_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RSVDB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// host-pointer mirrors of the reference API
// ---------------------------------------------------------------------------------------------------------------------
int rsvdb_generate_omega_host(rsvdb_ctx* c, int64_t n, int l, uint64_t seed, double* Omega, int64_t ldo) {
  if (!c || !Omega || n < 0 || l < 0 || ldo < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "generate_omega: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  RSVDB_CUDA(c, c->io_ws.reserve(IoArena::pad((size_t)n * l) * 8 + 256));
  IoArena ar(c); double* dO = ar.take((size_t)n * l);
  RSVDB_TRY(rsvdb_generate_omega_dev(c, n, l, seed, dO, n));
  RSVDB_TRY(d2h(c, Omega, ldo, dO, n, n, l));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_intermediate_step_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                                 int l, int q, double* Q, int64_t ldq) {
  if (!c || !A || !Omega || !Q || m < 0 || n <= 0 || l <= 0 || q < 0 || lda < m || ldo < n || ldq < m)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "intermediate_step: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)ldA * l)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dO = ar.take((size_t)ldO * l); double* dQ = ar.take((size_t)ldA * l);
  RSVDB_TRY(h2d(c, dO, ldO, Omega, ldo, n, l));
  const HostUpload up{A, lda};
  RSVDB_TRY(range_finder(c, dA, m, n, ldA, dO, ldO, l, q, dQ, ldA, &up));
  RSVDB_TRY(d2h(c, Q, ldq, dQ, ldA, m, l));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_rsvd_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega, int64_t ldo, uint64_t seed,
                    int l, int q, int method, double* U, int64_t ldu, double* S, double* V, int64_t ldv) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (method != 0 && method != 1 && method != 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Unsupported SVD method");
  if (!A || !U || !S || !V || m < 0 || n <= 0 || l <= 0 || q < 0 || lda < m || (Omega && ldo < n) || ldu < m || ldv < n)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rSVD: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t k = std::min<int64_t>(l, n);
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)ldA * l) +
                       IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)l)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dO = ar.take((size_t)ldO * l); double* dU = ar.take((size_t)ldA * l);
  double* dV = ar.take((size_t)ldO * l); double* dS = ar.take((size_t)l);
  if (Omega) { RSVDB_TRY(h2d(c, dO, ldO, Omega, ldo, n, l)); }
  else { RSVDB_TRY(rsvdb_generate_omega_dev(c, n, l, seed, dO, ldO)); }
  const HostUpload up{A, lda};                          // A is uploaded block by block underneath the first product
  RSVDB_TRY(rsvd_device(c, dA, m, n, ldA, dO, ldO, l, q, method, dU, ldA, dS, dV, ldO, seed, &up));
  RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, k));
  RSVDB_TRY(d2h(c, S, k, dS, k, k, 1));
  RSVDB_TRY(d2h(c, V, ldv, dV, ldO, n, k));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  RSVDB_TRY(check_svd_converged(c));
  return RSVDB_OK;
}

}  // extern "C"

namespace {
// PCA<method>::initialize() pre-pass on the freshly uploaded matrix (PCA/include/PCA_class.hpp:30-41)
struct PcaPre { int normalize; double* mean; double* stddev; };
int pca_prepass(rsvdb_ctx* c, double* dA, int64_t m, int64_t n, int64_t ldA, double* dMean, double* dSd, const PcaPre& pre) {
  RSVDB_TRY(column_stats(c, dA, m, n, ldA, dMean, pre.normalize ? dSd : nullptr, nullptr));
  RSVDB_TRY(center_columns(c, dA, m, n, ldA, dMean, pre.normalize ? dSd : nullptr));
  RSVDB_TRY(d2h(c, pre.mean, n, dMean, n, n, 1));
  if (pre.normalize && pre.stddev) RSVDB_TRY(d2h(c, pre.stddev, n, dSd, n, n, 1));
  return 0;
}

int svd_host_impl(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, int method, int r, uint64_t seed, double* U,
                  int64_t ldu, double* S, double* V, int64_t ldv, int* found, const PcaPre* pre) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (method != 0 && method != 1 && method != 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Unsupported SVD method");
  if (!A || !U || !S || !V || m <= 0 || n <= 0 || lda < m || ldu < m || ldv < n || r < 0)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "SVD: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t k = std::min(m, n);
  const int64_t ldA = even_ld(m), ldN = even_ld(n);
  if (method == RSVDB_SVD_POWER) {
    const int dim = r ? r : (int)k;
    if (dim > k) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "SVD<Power>: r larger than min(rows, cols)");
    const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldN * m) + IoArena::pad((size_t)ldA * m) +
                         IoArena::pad((size_t)ldN * dim) + IoArena::pad((size_t)k) + 2 * IoArena::pad((size_t)n)) * 8 + 256;
    RSVDB_CUDA(c, c->io_ws.reserve(need));
    IoArena ar(c);
    double* dA = ar.take((size_t)ldA * n); double* dAt = ar.take((size_t)ldN * m); double* dU = ar.take((size_t)ldA * m);
    double* dV = ar.take((size_t)ldN * dim); double* dS = ar.take((size_t)k);
    double* dMean = ar.take((size_t)n); double* dSd = ar.take((size_t)n);
    RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
    if (pre) RSVDB_TRY(pca_prepass(c, dA, m, n, ldA, dMean, dSd, *pre));
    RSVDB_TRY(transpose2d(c, dA, ldA, dAt, ldN, m, n));
    int f = 0;
    RSVDB_TRY(small_svd_power_t(c, dAt, ldN, m, n, r, seed, dU, ldA, (int)m, dS, dV, ldN, &f));
    RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, m));
    RSVDB_TRY(d2h(c, S, k, dS, k, k, 1));
    RSVDB_TRY(d2h(c, V, ldv, dV, ldN, n, dim));
    RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (found) *found = f;
    return RSVDB_OK;
  }
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldA * k) + IoArena::pad((size_t)ldN * k) +
                       IoArena::pad((size_t)k) + 2 * IoArena::pad((size_t)n)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dU = ar.take((size_t)ldA * k); double* dV = ar.take((size_t)ldN * k);
  double* dS = ar.take((size_t)k);
  double* dMean = ar.take((size_t)n); double* dSd = ar.take((size_t)n);
  RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
  if (pre) RSVDB_TRY(pca_prepass(c, dA, m, n, ldA, dMean, dSd, *pre));
  RSVDB_TRY(small_svd_jacobi(c, dA, ldA, nullptr, 0, m, n, dU, ldA, dS, dV, ldN));
  RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, k));
  RSVDB_TRY(d2h(c, S, k, dS, k, k, 1));
  RSVDB_TRY(d2h(c, V, ldv, dV, ldN, n, k));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  RSVDB_TRY(check_svd_converged(c));
  if (found) *found = (int)k;
  return RSVDB_OK;
}

}  // namespace

extern "C" {

int rsvdb_svd_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, int method, int r, uint64_t seed, double* U,
                   int64_t ldu, double* S, double* V, int64_t ldv, int* found) {
  return svd_host_impl(c, A, m, n, lda, method, r, seed, U, ldu, S, V, ldv, found, nullptr);
}

int rsvdb_pca_host(rsvdb_ctx* c, const double* data, int64_t m, int64_t n, int64_t ld, int normalize, int method, int r, uint64_t seed,
                   double* mean, double* stddev, double* U, int64_t ldu, double* S, double* V, int64_t ldv, int* found) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (m < 2 || n < 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "PCA requires at least 2 rows and 2 columns.");   // PCA_class.hpp:50-54
  if (!mean) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "PCA: mean output is required");
  const PcaPre pre{normalize, mean, stddev};
  return svd_host_impl(c, data, m, n, ld, method, r, seed, U, ldu, S, V, ldv, found, &pre);
}

int rsvdb_column_stats_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, double* d_mean, double* d_stddev) {
  if (!c || !dA || !d_mean || m <= 0 || n <= 0 || lda < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "column_stats: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return column_stats(c, dA, m, n, lda, d_mean, d_stddev, nullptr);
}

int rsvdb_center_columns_dev(rsvdb_ctx* c, double* dA, int64_t m, int64_t n, int64_t lda, const double* d_mean, const double* d_stddev) {
  if (!c || !dA || !d_mean || m < 0 || n <= 0 || lda < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "center_columns: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return center_columns(c, dA, m, n, lda, d_mean, d_stddev);
}

int rsvdb_rpca_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* d_mean, const double* d_stddev,
                   const double* dOmega, int64_t ldo, int l, int q, int method, double* dU, int64_t ldu, double* dS, double* dV,
                   int64_t ldv) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (method != 0 && method != 1 && method != 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Unsupported SVD method");
  if (!dA || !d_mean || !dOmega || !dU || !dS || !dV || m < 0 || n <= 0 || l <= 0 || q < 0 || lda < m || ldo < n || ldu < m || ldv < n)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rPCA: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  RSVDB_CUDA(c, c->pca_ws.reserve(PcaScratch::total(n, l) * sizeof(double)));
  Centering cen; cen.mu = d_mean;
  if (d_stddev) {
    double* inv = c->pca_ws.ptr + PcaScratch::inv_sd_off(n);
    k_reciprocal<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_stddev, inv, n);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
    cen.inv_sd = inv;
  }
  return rsvd_device(c, dA, m, n, lda, dOmega, ldo, l, q, method, dU, ldu, dS, dV, ldv, 0, nullptr, &cen);
}

int rsvdb_rpca_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, int normalize, const double* Omega, int64_t ldo,
                    uint64_t seed, int l, int q, int method, double* mean, double* stddev, double* U, int64_t ldu, double* S, double* V,
                    int64_t ldv) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (method != 0 && method != 1 && method != 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Unsupported SVD method");
  if (m < 2 || n < 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "PCA requires at least 2 rows and 2 columns.");
  if (!A || !U || !S || !V || !mean || l <= 0 || q < 0 || lda < m || (Omega && ldo < n) || ldu < m || ldv < n)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rPCA: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t k = std::min<int64_t>(l, n);
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)ldA * l) +
                       IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)l) + 2 * IoArena::pad((size_t)n)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  RSVDB_CUDA(c, c->pca_ws.reserve(PcaScratch::total(n, l) * sizeof(double)));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dO = ar.take((size_t)ldO * l); double* dU = ar.take((size_t)ldA * l);
  double* dV = ar.take((size_t)ldO * l); double* dS = ar.take((size_t)l);
  double* dMean = ar.take((size_t)n); double* dSd = ar.take((size_t)n);
  RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
  if (Omega) { RSVDB_TRY(h2d(c, dO, ldO, Omega, ldo, n, l)); }
  else { RSVDB_TRY(rsvdb_generate_omega_dev(c, n, l, seed, dO, ldO)); }
  double* inv = normalize ? c->pca_ws.ptr + PcaScratch::inv_sd_off(n) : nullptr;
  RSVDB_TRY(column_stats(c, dA, m, n, ldA, dMean, normalize ? dSd : nullptr, inv));
  Centering cen; cen.mu = dMean; cen.inv_sd = inv;
  RSVDB_TRY(rsvd_device(c, dA, m, n, ldA, dO, ldO, l, q, method, dU, ldA, dS, dV, ldO, seed, nullptr, &cen));
  RSVDB_TRY(d2h(c, mean, n, dMean, n, n, 1));
  if (normalize && stddev) RSVDB_TRY(d2h(c, stddev, n, dSd, n, n, 1));
  RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, k));
  RSVDB_TRY(d2h(c, S, k, dS, k, k, 1));
  RSVDB_TRY(d2h(c, V, ldv, dV, ldO, n, k));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  RSVDB_TRY(check_svd_converged(c));
  return RSVDB_OK;
}

int rsvdb_pca_project_host(rsvdb_ctx* c, const double* data, int64_t r, int64_t n, int64_t ld, const double* mean, const double* V,
                           int64_t ldv, int k, double* out, int64_t ldout) {
  if (!c || !data || !mean || !V || !out || r <= 0 || n <= 0 || k <= 0 || ld < r || ldv < n || ldout < r)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "projectToPCA: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldR = even_ld(r), ldN = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldR * n) + IoArena::pad((size_t)ldN * k) + IoArena::pad((size_t)ldR * k) + IoArena::pad((size_t)n) +
                       IoArena::pad((size_t)k)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dD = ar.take((size_t)ldR * n); double* dV = ar.take((size_t)ldN * k); double* dOut = ar.take((size_t)ldR * k);
  double* dMean = ar.take((size_t)n); double* dw = ar.take((size_t)k);
  RSVDB_TRY(h2d(c, dD, ldR, data, ld, r, n));
  RSVDB_TRY(h2d(c, dV, ldN, V, ldv, n, k));
  RSVDB_TRY(h2d(c, dMean, n, mean, n, n, 1));
  int nl = 0;                                                    // (data - 1 mean^T) V = data V - 1 (mean^T V)
  RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, dD, r, n, ldR, dV, ldN, k, dOut, ldR, &nl));
  c->launches += nl;
  RSVDB_TRY(weighted_colsum(c, dV, ldN, n, k, dMean, dw));
  RSVDB_TRY(sub_col_const(c, dOut, ldR, r, k, dw));
  RSVDB_TRY(d2h(c, out, ldout, dOut, ldR, r, k));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_pca_reconstruct_host(rsvdb_ctx* c, const double* pc, int64_t r, int k, int64_t ldp, const double* mean, const double* V,
                               int64_t ldv, int64_t n, double* out, int64_t ldout) {
  if (!c || !pc || !mean || !V || !out || r <= 0 || n <= 0 || k <= 0 || ldp < r || ldv < n || ldout < r)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "reconstructFromPCA: bad argument");
  if (n > INT32_MAX) return fail(c, RSVDB_ERR_UNSUPPORTED, "reconstructFromPCA: more than 2^31-1 features");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldR = even_ld(r), ldN = even_ld(n), ldK = even_ld(k);
  const size_t need = (IoArena::pad((size_t)ldR * k) + IoArena::pad((size_t)ldN * k) + IoArena::pad((size_t)ldK * n) + IoArena::pad((size_t)ldR * n) +
                       IoArena::pad((size_t)n)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dP = ar.take((size_t)ldR * k); double* dV = ar.take((size_t)ldN * k); double* dVt = ar.take((size_t)ldK * n);
  double* dOut = ar.take((size_t)ldR * n); double* dMean = ar.take((size_t)n);
  RSVDB_TRY(h2d(c, dP, ldR, pc, ldp, r, k));
  RSVDB_TRY(h2d(c, dV, ldN, V, ldv, n, k));
  RSVDB_TRY(h2d(c, dMean, n, mean, n, n, 1));
  RSVDB_TRY(transpose2d(c, dV, ldN, dVt, ldK, n, k));           // V^T (k x n) as the right operand
  int nl = 0;
  RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, dP, r, k, ldR, dVt, ldK, (int)n, dOut, ldR, &nl));
  c->launches += nl;
  RSVDB_TRY(add_row_vector(c, dOut, ldR, r, n, dMean, 1.0));
  RSVDB_TRY(d2h(c, out, ldout, dOut, ldR, r, n));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_pod_shape(int variant, int64_t Nh, int64_t ns, int r, int svd_type, int64_t* w_cols_full, int64_t* sigma_len) {
  PodShape sh;
  if (variant < 0 || variant > 3 || Nh <= 0 || ns <= 0 || r <= 0 || !pod_shape(variant, Nh, ns, r, svd_type, &sh)) return RSVDB_ERR_INVALID_ARGUMENT;
  if (w_cols_full) *w_cols_full = sh.w_cols_full;
  if (sigma_len) *sigma_len = sh.sigma_len;
  return RSVDB_OK;
}

int rsvdb_pod_dev(rsvdb_ctx* c, int variant, const double* dS, int64_t Nh, int64_t ns, int64_t lds, const double* dXh, int64_t ldx,
                  const double* dD, int64_t ldd, int r, double tol, int svd_type, uint64_t seed, const double* dOmega, int64_t ldo,
                  double* dW, int64_t ldw, double* d_sigma, int* N) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (!dS || !dW || !d_sigma || !N || lds < Nh || ldw < Nh) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "POD: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return pod_device(c, variant, dS, Nh, ns, lds, dXh, ldx, dD, ldd, r, tol, svd_type, seed, dOmega, ldo, dW, ldw, d_sigma, N);
}

int rsvdb_pod_host(rsvdb_ctx* c, int variant, const double* S, int64_t Nh, int64_t ns, int64_t lds, const double* Xh, int64_t ldx,
                   const double* D, int64_t ldd, int r, double tol, int svd_type, uint64_t seed, const double* Omega, int64_t ldo,
                   double* W, int64_t ldw, double* sigma, int* N) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  PodShape sh;
  if (variant < 0 || variant > 3 || Nh <= 0 || ns <= 0 || r <= 0) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "POD: bad argument");
  if (!pod_shape(variant, Nh, ns, r, svd_type, &sh))
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "The svd_type should be in [0,5]. Check 'svd_type' in the parameter file.");
  if (!S || !W || !sigma || !N || lds < Nh || ldw < Nh || (variant >= 2 && (!Xh || ldx < Nh)) || (variant == 3 && (!D || ldd < ns)))
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "POD: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldS = even_ld(Nh), ldD = even_ld(ns);
  const int64_t ob = variant == 0 ? ns : std::min(ns, Nh), ldOm = even_ld(ob);       // rows of Omega
  if (Omega && ldo < ob) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "POD: Omega leading dimension too small");
  const size_t need = (IoArena::pad((size_t)ldS * ns) + IoArena::pad((size_t)ldS * sh.w_cols_full) + IoArena::pad((size_t)sh.sigma_len) +
                       (variant >= 2 ? IoArena::pad((size_t)ldS * Nh) : 0) + (variant == 3 ? IoArena::pad((size_t)ldD * ns) : 0) +
                       (Omega ? IoArena::pad((size_t)ldOm * r) : 0)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dS = ar.take((size_t)ldS * ns); double* dW = ar.take((size_t)ldS * sh.w_cols_full); double* dSig = ar.take((size_t)sh.sigma_len);
  double* dX = variant >= 2 ? ar.take((size_t)ldS * Nh) : nullptr; double* dD = variant == 3 ? ar.take((size_t)ldD * ns) : nullptr;
  double* dOm = Omega ? ar.take((size_t)ldOm * r) : nullptr;
  RSVDB_TRY(h2d(c, dS, ldS, S, lds, Nh, ns));
  if (dX) RSVDB_TRY(h2d(c, dX, ldS, Xh, ldx, Nh, Nh));
  if (dD) RSVDB_TRY(h2d(c, dD, ldD, D, ldd, ns, ns));
  if (dOm) RSVDB_TRY(h2d(c, dOm, ldOm, Omega, ldo, ob, r));
  RSVDB_TRY(pod_device(c, variant, dS, Nh, ns, ldS, dX, ldS, dD, ldD, r, tol, svd_type, seed, dOm, ldOm, dW, ldS, dSig, N));
  RSVDB_TRY(d2h(c, W, ldw, dW, ldS, Nh, *N));             // the basis after conservativeResize(NoChange, N), POD.cpp:221
  RSVDB_TRY(d2h(c, sigma, sh.sigma_len, dSig, sh.sigma_len, sh.sigma_len, 1));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  RSVDB_TRY(check_svd_converged(c));
  return RSVDB_OK;
}

int rsvdb_image_compress_host(rsvdb_ctx* c, const double* image, int64_t m, int64_t n, int64_t ld, int k, int normalize, const double* Omega,
                              int64_t ldo, uint64_t seed, double* original_min, double* original_max, double* U, int64_t ldu, double* S,
                              double* V, int64_t ldv, int* degree) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (!image || !U || !S || !V || m <= 0 || n <= 0 || ld < m || ldu < m || ldv < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Image::compress: bad argument");
  if (k == -1) k = (int)(std::min(m, n) / 4);                       // image_com.cpp:293-295
  const int l = k + 10;                                             // oversampling p = 10, :297-298
  if (k < 0 || l > n || (Omega && ldo < n)) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Image::compress: k + 10 must not exceed the image width");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)ldA * l) +
                       IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)l) + IoArena::pad(8)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dO = ar.take((size_t)ldO * l); double* dU = ar.take((size_t)ldA * l);
  double* dV = ar.take((size_t)ldO * l); double* dS = ar.take((size_t)l); double* dmm = ar.take(8);
  RSVDB_TRY(h2d(c, dA, ldA, image, ld, m, n));
  if (Omega) { RSVDB_TRY(h2d(c, dO, ldO, Omega, ldo, n, l)); }
  else { RSVDB_TRY(rsvdb_generate_omega_dev(c, n, l, seed, dO, ldO)); }
  RSVDB_TRY(image_minmax(c, dA, m, n, ldA, dmm));
  if (normalize) RSVDB_TRY(image_affine(c, dA, m, n, ldA, dmm, false));
  RSVDB_TRY(rsvd_device(c, dA, m, n, ldA, dO, ldO, l, /*q=*/1, RSVDB_SVD_POWER, dU, ldA, dS, dV, ldO, seed));
  double mm[2] = {0.0, 0.0};
  RSVDB_TRY(d2h(c, mm, 2, dmm, 2, 2, 1));
  RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, l));
  RSVDB_TRY(d2h(c, S, l, dS, l, l, 1));
  RSVDB_TRY(d2h(c, V, ldv, dV, ldO, n, l));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (original_min) *original_min = mm[0];
  if (original_max) *original_max = mm[1];
  if (degree) *degree = l;
  return RSVDB_OK;
}

int rsvdb_image_normalize_host(rsvdb_ctx* c, double* image, int64_t m, int64_t n, int64_t ld, int inverse, double* original_min,
                               double* original_max) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (!image || !original_min || !original_max || m <= 0 || n <= 0 || ld < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Image::normalize: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m);
  RSVDB_CUDA(c, c->io_ws.reserve((IoArena::pad((size_t)ldA * n) + IoArena::pad(8)) * 8 + 256));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dmm = ar.take(8);
  RSVDB_TRY(h2d(c, dA, ldA, image, ld, m, n));
  double mm[2] = {*original_min, *original_max};
  if (inverse) { RSVDB_TRY(h2d(c, dmm, 2, mm, 2, 2, 1)); }
  else { RSVDB_TRY(image_minmax(c, dA, m, n, ldA, dmm)); RSVDB_TRY(d2h(c, mm, 2, dmm, 2, 2, 1)); }
  RSVDB_TRY(image_affine(c, dA, m, n, ldA, dmm, inverse != 0));
  RSVDB_TRY(d2h(c, image, ld, dA, ldA, m, n));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  *original_min = mm[0]; *original_max = mm[1];
  return RSVDB_OK;
}

int rsvdb_image_reconstruct_host(rsvdb_ctx* c, const double* U, int64_t m, int64_t ldu, const double* S, const double* V, int64_t n, int64_t ldv,
                                 int l, int denormalize, double original_min, double original_max, double* out, int64_t ldout) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (!U || !S || !V || !out || m <= 0 || n <= 0 || l <= 0 || ldu < m || ldv < n || ldout < m || n > INT32_MAX)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Image::reconstruct: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldA * l) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)l) +
                       IoArena::pad(8)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dOut = ar.take((size_t)ldA * n); double* dU = ar.take((size_t)ldA * l); double* dV = ar.take((size_t)ldO * l);
  double* dS = ar.take((size_t)l); double* dmm = ar.take(8);
  RSVDB_TRY(h2d(c, dU, ldA, U, ldu, m, l));
  RSVDB_TRY(h2d(c, dV, ldO, V, ldv, n, l));
  RSVDB_TRY(h2d(c, dS, l, S, l, l, 1));
  const double mm[2] = {original_min, original_max};
  if (denormalize) RSVDB_TRY(h2d(c, dmm, 2, mm, 2, 2, 1));
  RSVDB_TRY(image_reconstruct(c, dU, m, ldA, dS, dV, n, ldO, l, denormalize ? dmm : nullptr, dOut, ldA));
  RSVDB_TRY(d2h(c, out, ldout, dOut, ldA, m, n));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_qr_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, int full, double* Q, int64_t ldq, double* R,
                  int64_t ldr) {
  if (!c || !A || !Q || !R || m <= 0 || n <= 0 || lda < m || ldq < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "QR: bad argument");
  if (!full && m < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "reduced QR needs rows >= cols (reference src/QR.cpp:78-79)");
  if (ldr < (full ? m : n)) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "QR: ldr too small");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m);
  if (!full) {
    const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)n * n)) * 8 + 256;
    RSVDB_CUDA(c, c->io_ws.reserve(need));
    IoArena ar(c); double* dA = ar.take((size_t)ldA * n); double* dR = ar.take((size_t)n * n);
    RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
    RSVDB_TRY(rsvdb_qr_dev(c, dA, m, (int)n, ldA, 0, dR));
    k_sign_normalise<<<(unsigned)n, 256, 0, c->stream>>>(dA, ldA, m, dR, n, (int)n, (int)n);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
    RSVDB_TRY(d2h(c, Q, ldq, dA, ldA, m, n));
    RSVDB_TRY(d2h(c, R, ldr, dR, n, n, n));
    RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
    return RSVDB_OK;
  }
  // full: Q m x m, R m x n
  const int64_t kk = std::min(m, n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)n * n) + IoArena::pad((size_t)ldA * m) +
                       IoArena::pad((size_t)n) + IoArena::pad((size_t)ldA * n)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dRs = ar.take((size_t)n * n); double* dQ = ar.take((size_t)ldA * m);
  double* dtau = ar.take((size_t)n); double* dRf = ar.take((size_t)ldA * n);
  RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
  int nl = 0;
  RSVDB_CUDA(c, house_full_qr(c->stream, dA, ldA, m, (int)n, dtau, dRs, dQ, ldA, &nl));
  c->launches += nl;
  // R (m x n) = [Rs (first min(m,n) rows); 0]
  RSVDB_CUDA(c, cudaMemsetAsync(dRf, 0, (size_t)ldA * n * 8, c->stream));
  RSVDB_TRY(copy2d(c, dRs, n, dRf, ldA, kk, (int)n));
  k_sign_normalise<<<(unsigned)kk, 256, 0, c->stream>>>(dQ, ldA, m, dRf, ldA, (int)n, (int)kk);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  RSVDB_TRY(d2h(c, Q, ldq, dQ, ldA, m, m));
  RSVDB_TRY(d2h(c, R, ldr, dRf, ldA, m, n));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_pm_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, uint64_t seed, double* sigma, double* u, double* v) {
  if (!c || !A || !sigma || !u || !v || m <= 0 || n <= 0 || lda < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "PM: bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m), ldN = even_ld(n);
  const size_t need = (IoArena::pad((size_t)ldA * n) + IoArena::pad((size_t)ldN * m) + IoArena::pad((size_t)ldA) + IoArena::pad((size_t)ldN) +
                       IoArena::pad((size_t)std::min(m, n))) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * n); double* dAt = ar.take((size_t)ldN * m); double* du = ar.take((size_t)ldA);
  double* dv = ar.take((size_t)ldN); double* dS = ar.take((size_t)std::min(m, n));
  RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, n));
  RSVDB_TRY(transpose2d(c, dA, ldA, dAt, ldN, m, n));
  int f = 0;
  RSVDB_TRY(small_svd_power_t(c, dAt, ldN, m, n, 1, seed, du, ldA, 1, dS, dv, ldN, &f));
  RSVDB_TRY(d2h(c, u, m, du, ldA, m, 1));
  RSVDB_TRY(d2h(c, v, n, dv, ldN, n, 1));
  RSVDB_TRY(d2h(c, sigma, 1, dS, 1, 1, 1));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

int rsvdb_csr_spmm_dev(rsvdb_ctx* c, int64_t m, const int64_t* rp, const int32_t* ci, const double* v, const double* X, int l, double* Y) {
  if (!c || m < 0 || l <= 0) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "csr_spmm: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return csr_spmm_rm(c, m, rp, ci, v, X, l, Y);
}

int rsvdb_rsvd_csr_dev(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rp, const int32_t* ci, const double* v,
                       const double* dOmega, int64_t ldo, uint64_t seed, int l, int q, int method, double* dU, int64_t ldu, double* dS,
                       double* dV, int64_t ldv) {
  if (!c || m < 0 || n <= 0 || nnz < 0 || ldo < n || ldu < m || ldv < n) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rsvd_csr: bad shape");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  return rsvd_csr_device(c, m, n, nnz, rp, ci, v, dOmega, ldo, l, q, method, dU, ldu, dS, dV, ldv, seed);
}

int rsvdb_rsvd_csr_host(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx, const double* values,
                        const double* Omega, int64_t ldo, uint64_t seed, int l, int q, int method, double* U, int64_t ldu, double* S,
                        double* V, int64_t ldv) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (method != 0 && method != 1 && method != 2) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Unsupported SVD method");
  if (!rowptr || (nnz > 0 && (!colidx || !values)) || !U || !S || !V || m < 0 || n <= 0 || nnz < 0 || l <= 0 || q < 0 || (Omega && ldo < n) ||
      ldu < m || ldv < n)
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "rSVD (CSR): bad argument");
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t k = std::min<int64_t>(l, n);
  const int64_t ldA = even_ld(m), ldO = even_ld(n);
  const size_t need = (IoArena::pad((size_t)m + 2) + IoArena::pad((size_t)nnz / 2 + 2) + IoArena::pad((size_t)nnz + 2) + IoArena::pad((size_t)ldO * l) +
                       IoArena::pad((size_t)ldA * l) + IoArena::pad((size_t)ldO * l) + IoArena::pad((size_t)l)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  int64_t* drp = reinterpret_cast<int64_t*>(ar.take((size_t)m + 2)); int32_t* dci = reinterpret_cast<int32_t*>(ar.take((size_t)nnz / 2 + 2));
  double* dv = ar.take((size_t)nnz + 2); double* dO = ar.take((size_t)ldO * l); double* dU = ar.take((size_t)ldA * l);
  double* dV = ar.take((size_t)ldO * l); double* dS = ar.take((size_t)l);
  {
    PhaseTimer pt(c, PH_COPY);
    RSVDB_CUDA(c, cudaMemcpyAsync(drp, rowptr, (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) {
      RSVDB_CUDA(c, cudaMemcpyAsync(dci, colidx, (size_t)nnz * 4, cudaMemcpyHostToDevice, c->stream));
      RSVDB_CUDA(c, cudaMemcpyAsync(dv, values, (size_t)nnz * 8, cudaMemcpyHostToDevice, c->stream));
    }
  }
  if (Omega) { RSVDB_TRY(h2d(c, dO, ldO, Omega, ldo, n, l)); }
  else { RSVDB_TRY(rsvdb_generate_omega_dev(c, n, l, seed, dO, ldO)); }
  RSVDB_TRY(rsvd_csr_device(c, m, n, nnz, drp, dci, dv, dO, ldO, l, q, method, dU, ldA, dS, dV, ldO, seed));
  RSVDB_TRY(d2h(c, U, ldu, dU, ldA, m, k));
  RSVDB_TRY(d2h(c, S, k, dS, k, k, 1));
  RSVDB_TRY(d2h(c, V, ldv, dV, ldO, n, k));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  RSVDB_TRY(check_svd_converged(c));
  return RSVDB_OK;
}

int rsvdb_gemm_host(rsvdb_ctx* c, const double* A, int64_t m, int64_t ka, int64_t lda, const double* B, int64_t kb, int64_t n,
                    int64_t ldb, double* C, int64_t ldc) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  if (ka != kb) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "Matrices dimensions are not compatible for manual matrix multiplication");
  if (!A || !B || !C || m < 0 || n < 0 || ka < 0 || lda < m || ldb < kb || ldc < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "gemm: bad argument");
  if (m == 0 || n == 0) return RSVDB_OK;
  RSVDB_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldA = even_ld(m), ldB = even_ld(ka);
  const size_t need = (IoArena::pad((size_t)ldA * std::max<int64_t>(ka, 1)) + IoArena::pad((size_t)ldB * n) + IoArena::pad((size_t)ldA * n)) * 8 + 256;
  RSVDB_CUDA(c, c->io_ws.reserve(need));
  IoArena ar(c);
  double* dA = ar.take((size_t)ldA * std::max<int64_t>(ka, 1)); double* dB = ar.take((size_t)ldB * n); double* dC = ar.take((size_t)ldA * n);
  RSVDB_TRY(h2d(c, dA, ldA, A, lda, m, ka));
  RSVDB_TRY(h2d(c, dB, ldB, B, ldb, ka, n));
  for (int64_t n0 = 0; n0 < n; n0 += 128) {   // the skinny kernel takes <= 128 columns per launch; wider B goes through in slabs
    const int nc = (int)std::min<int64_t>(128, n - n0);
    RSVDB_TRY(rsvdb_gemm_an_dev(c, dA, m, ka, ldA, dB + (size_t)n0 * ldB, ldB, nc, dC + (size_t)n0 * ldA, ldA));
  }
  RSVDB_TRY(d2h(c, C, ldc, dC, ldA, m, n));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

}  // extern "C"
