// Device-side orchestration of the rSVD hot path: range finder, projection, small SVD, back-projection.
// Mirrors reference src/rSVD.cpp:57-133 step by step; every product and factorisation is one of the sm_100a kernels
// of this directory, and with c->nranks > 1 the row-sharded variant exchanges only A^T Q partial sums and TSQR R factors.
#include "pipeline.cuh"

#include <algorithm>

#include "comm.cuh"
#include "jacobi.cuh"
#include "tsqr.cuh"

namespace rsvdb {

namespace {

// gathered[p][k][i] (each rank's l x l block, ld = l) -> stack[(p*l + i) + k * (P*l)]
__global__ void k_restack(const double* __restrict__ gathered, double* __restrict__ stack, int l, int P) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = P * l * l;
  if (e < total) {
    const int i = e % l, k = (e / l) % l, p = e / (l * l);
    stack[(size_t)k * P * l + (size_t)p * l + i] = gathered[e];
  }
}

__global__ void k_transpose(const double* __restrict__ src, long long lds, double* __restrict__ dst, long long ldd, long long rows, long long cols,
                            long long cblock0) {
  // dst (cols x rows) = src (rows x cols)^T; grid.y covers column blocks [cblock0, cblock0 + gridDim.y)
  __shared__ double tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32, c0 = (cblock0 + blockIdx.y) * 32;
  for (int cc = threadIdx.y; cc < 32; cc += 8) {
    const long long r = r0 + threadIdx.x, c = c0 + cc;
    tile[cc][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)c * lds + r] : 0.0;
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += 8) {
    const long long r = r0 + rr, c = c0 + threadIdx.x;
    if (r < rows && c < cols) dst[(size_t)r * ldd + c] = tile[threadIdx.x][rr];
  }
}

__global__ void k_copy2d(const double* __restrict__ src, long long lds, double* __restrict__ dst, long long ldd, long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) for (int k = blockIdx.y; k < cols; k += gridDim.y) dst[(size_t)k * ldd + i] = src[(size_t)k * lds + i];
}

}  // namespace

int transpose2d(rsvdb_ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t cblocks = (cols + 31) / 32;
  for (int64_t cb = 0; cb < cblocks; cb += 65535) {                      // gridDim.y is limited to 65535
    dim3 g((unsigned)((rows + 31) / 32), (unsigned)std::min<int64_t>(65535, cblocks - cb));
    k_transpose<<<g, dim3(32, 8), 0, c->stream>>>(src, lds, dst, ldd, rows, cols, cb);
    RSVDB_CUDA(c, cudaGetLastError());
    ++c->launches;
  }
  return 0;
}

int copy2d(rsvdb_ctx* c, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 g((unsigned)((rows + 255) / 256), (unsigned)std::min(cols, 128));
  k_copy2d<<<g, 256, 0, c->stream>>>(src, lds, dst, ldd, rows, cols);
  RSVDB_CUDA(c, cudaGetLastError());
  ++c->launches;
  return 0;
}

namespace {
constexpr int QR_FAST_MAX = 100;    // widest panel the blocked / cluster TSQR holds in shared memory
constexpr int QR_WIDE_BLOCK = 96;

// Y (rows x cols, ldy) -= P (rows x cols, ldp)
__global__ void k_sub(double* __restrict__ Y, long long ldy, const double* __restrict__ P, long long ldp, long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) for (int k = blockIdx.y; k < cols; k += gridDim.y) Y[(size_t)k * ldy + i] -= P[(size_t)k * ldp + i];
}
// R[r0 + i, c0 + j] (+)= W[i, j]
__global__ void k_place(double* __restrict__ R, int ldr, int r0, int c0, const double* __restrict__ Wm, int ldw, int rows, int cols, int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < rows * cols) {
    const int i = e % rows, j = e / rows;
    double* dst = R + (size_t)(c0 + j) * ldr + r0 + i;
    *dst = accumulate ? *dst + Wm[(size_t)j * ldw + i] : Wm[(size_t)j * ldw + i];
  }
}
}  // namespace

// Panels wider than the shared-memory TSQR (l > 100): block Gram-Schmidt between column blocks of <= 96, Householder TSQR
// inside each block.  Block k is projected twice against the finished blocks ("twice is enough"), factored, projected
// once more and factored again, so that Q stays orthonormal to working precision even when a block is numerically
// dependent on its predecessors (rank-deficient sketches); R is assembled from the projection coefficients.
// All products run on the DMMA GEMM kernels; with row shards the coefficients are all-reduced.
static int qr_wide(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool sharded, const double** Rout) {
  // a block (<= 104 columns, projected against its predecessors) is factored like any sketch of the pipeline: guarded CholeskyQR2
  // when the caller is the pipeline's orthonormalize (c->wide_fast), Householder TSQR for the QR class
  auto block_qr = [&](double* Yk, int ck, const double** Rk) -> int {
    return c->wide_fast ? orthonormalize(c, Yk, rows, ck, ldy, sharded, Rk) : qr_inplace(c, Yk, rows, ck, ldy, sharded, Rk);
  };
  const int nblk = (l + QR_WIDE_BLOCK - 1) / QR_WIDE_BLOCK;
  const int bw = (((l + nblk - 1) / nblk) + 7) & ~7;
  const size_t d_R = (size_t)l * l, d_W = (size_t)l * bw, d_P = (size_t)std::max<int64_t>(rows, 1) * bw, d_s = (size_t)bw * bw;
  RSVDB_CUDA(c, c->wide_ws.reserve((d_R + 2 * d_W + d_P + 3 * d_s + 64) * sizeof(double)));
  double* R = c->wide_ws.ptr; double* Wb = R + d_R; double* W2 = Wb + d_W; double* P = W2 + d_W;
  double* R1 = P + d_P; double* R2 = R1 + d_s; double* R21 = R2 + d_s;
  cudaStream_t st = c->stream;
  RSVDB_CUDA(c, cudaMemsetAsync(R, 0, d_R * sizeof(double), st));
  const bool dist = sharded && c->nranks > 1;
  int nl = 0;
  auto project = [&](double* Yk, int col0, int ck, double* Wout) -> int {       // Wout = Qp^T Yk ; Yk -= Qp Wout
    RSVDB_CUDA(c, gemm_at(c->gemm_ws, st, c->nsm, Y, rows, col0, ldy, Yk, ldy, ck, Wout, col0, 0, &nl));
    if (dist) RSVDB_TRY(comm_allreduce_sum(c, Wout, (size_t)col0 * ck));
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, st, c->nsm, Y, rows, col0, ldy, Wout, col0, ck, P, rows, &nl));
    if (rows > 0) { dim3 g((unsigned)((rows + 255) / 256), (unsigned)std::min(ck, 64)); k_sub<<<g, 256, 0, st>>>(Yk, ldy, P, rows, rows, ck); ++nl; }
    return 0;
  };
  for (int col0 = 0; col0 < l; col0 += bw) {
    const int ck = std::min(bw, l - col0);
    double* Yk = Y + (size_t)col0 * ldy;
    const double* Rk = nullptr;
    if (col0 == 0) {
      RSVDB_TRY(block_qr(Yk, ck, &Rk));
      k_place<<<(ck * ck + 255) / 256, 256, 0, st>>>(R, l, 0, 0, Rk, ck, ck, ck, 0); ++nl;
      continue;
    }
    // Y_k = Qp (W_a + W_b) + Y_k''            (two projections)
    RSVDB_TRY(project(Yk, col0, ck, Wb));
    k_place<<<(col0 * ck + 255) / 256, 256, 0, st>>>(R, l, 0, col0, Wb, col0, col0, ck, 0); ++nl;
    RSVDB_TRY(project(Yk, col0, ck, Wb));
    k_place<<<(col0 * ck + 255) / 256, 256, 0, st>>>(R, l, 0, col0, Wb, col0, col0, ck, 1); ++nl;
    // Y_k'' = Q1 R1
    RSVDB_TRY(block_qr(Yk, ck, &Rk));
    RSVDB_CUDA(c, cudaMemcpyAsync(R1, Rk, (size_t)ck * ck * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // Q1 = Qp W_c + Q2 R2                     (re-orthogonalise the block against its predecessors)
    RSVDB_TRY(project(Yk, col0, ck, W2));
    RSVDB_TRY(block_qr(Yk, ck, &Rk));
    RSVDB_CUDA(c, cudaMemcpyAsync(R2, Rk, (size_t)ck * ck * sizeof(double), cudaMemcpyDeviceToDevice, st));
    // R[0:col0, blk] += W_c R1 ;  R[blk, blk] = R2 R1
    RSVDB_CUDA(c, gemm_generic(st, 0, 0, col0, ck, ck, 1.0, W2, col0, R1, ck, 1.0, R + (size_t)col0 * l, l)); ++nl;
    RSVDB_CUDA(c, gemm_generic(st, 0, 0, ck, ck, ck, 1.0, R2, ck, R1, ck, 0.0, R21, ck)); ++nl;
    k_place<<<(ck * ck + 255) / 256, 256, 0, st>>>(R, l, col0, col0, R21, ck, ck, ck, 0); ++nl;
  }
  RSVDB_CUDA(c, cudaGetLastError());
  c->launches += nl;
  if (Rout) *Rout = R;
  return 0;
}

int qr_inplace(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool sharded, const double** R) {
  if (l > QR_FAST_MAX) {
    // The wide path and the plain TSQR issue different collective sequences, so with row shards every rank must take
    // the same branch although shard heights differ by one row (reference split rule): the ranks agree on the SHORTEST
    // shard first (one 8-byte all-reduce; only panels wider than the shared-memory TSQR get here).
    int64_t rows_min = rows;
    if (sharded && c->nranks > 1) {
      RSVDB_CUDA(c, c->svd_ws.reserve(64));
      const double neg = -(double)rows;                                    // min(rows) = -max(-rows); exact in FP64
      RSVDB_CUDA(c, cudaMemcpyAsync(c->svd_ws.ptr, &neg, sizeof(double), cudaMemcpyHostToDevice, c->stream));
      RSVDB_TRY(comm_allreduce_max(c, c->svd_ws.ptr, 1));
      double got = 0.0;
      RSVDB_CUDA(c, cudaMemcpyAsync(&got, c->svd_ws.ptr, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
      rows_min = (int64_t)(-got);
    }
    if (rows_min >= 4 * (int64_t)l) return qr_wide(c, Y, rows, l, ldy, sharded, R);
  }
  PhaseTimer pt(c, PH_QR);
  int k = 0;
  Tsqr t(&c->qr_ws, c->side_stream, c->side_ev);
  RSVDB_CUDA(c, t.plan(rows, l));
  RSVDB_CUDA(c, t.factor(c->stream, Y, ldy, &k));
  const bool dist = sharded && c->nranks > 1;
  if (!dist) {
    RSVDB_CUDA(c, t.form_q(c->stream, Y, ldy, nullptr, 0, &k));
    if (R) *R = t.R_local();
    c->launches += k;
    return 0;
  }
  // TSQR over row shards: all-gather the l x l R factors, factor the (P*l) x l stack redundantly on every rank
  // (deterministic kernels => identical bits), and push this rank's l x l block of the stack's Q down the local tree.
  const int P = c->nranks;
  Tsqr t2(&c->qr2_ws);
  RSVDB_CUDA(c, t2.plan((long long)P * l, l));
  // scratch behind t2's own storage: gathered blocks + the stack
  const size_t extra = (size_t)2 * P * l * l * sizeof(double);
  GemmWorkspace& gw = c->svd_ws;
  RSVDB_CUDA(c, gw.reserve(extra));
  double* gathered = gw.ptr; double* stack = gw.ptr + (size_t)P * l * l;
  {
    PhaseTimer pc(c, PH_COMM);
    RSVDB_TRY(comm_allgather(c, t.R_local(), gathered, (size_t)l * l));
  }
  k_restack<<<(P * l * l + 255) / 256, 256, 0, c->stream>>>(gathered, stack, l, P);
  RSVDB_CUDA(c, cudaGetLastError()); ++k;
  RSVDB_CUDA(c, t2.factor(c->stream, stack, (long long)P * l, &k));
  RSVDB_CUDA(c, t2.form_q(c->stream, stack, (long long)P * l, nullptr, 0, &k));
  RSVDB_CUDA(c, t.form_q(c->stream, Y, ldy, stack + (size_t)c->rank * l, (long long)P * l, &k));
  if (R) *R = t2.R_local();
  c->launches += k;
  return 0;
}

static double* cen_xs(rsvdb_ctx* c, int64_t n) { return c->pca_ws.ptr + PcaScratch::xs_off(n); }   // layout: pca.cuh

static int gemm_an_phase(rsvdb_ctx* c, const double* A, int64_t M, int64_t K, int64_t lda, const double* X, int64_t ldx, int N,
                         double* Y, int64_t ldy, const Centering* cen = nullptr) {
  double* w = nullptr;
  if (cen) {                                                 // Ac X = A (D X) - 1 (mu^T D X)
    PhaseTimer po(c, PH_OTHER);
    double* Xs = cen_xs(c, K); w = Xs + PcaScratch::pad((size_t)K * N);
    if (cen->inv_sd) { RSVDB_TRY(scale_rows_copy(c, X, ldx, Xs, K, K, N, cen->inv_sd)); X = Xs; ldx = K; }
    RSVDB_TRY(weighted_colsum(c, X, ldx, K, N, cen->mu, w));
  }
  {
    PhaseTimer pt(c, PH_GEMM_AN);
    int k = 0;
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, A, M, K, lda, X, ldx, N, Y, ldy, &k));
    c->launches += k;
  }
  if (cen) { PhaseTimer po(c, PH_OTHER); RSVDB_TRY(sub_col_const(c, Y, ldy, M, N, w)); }
  return 0;
}
static int gemm_at_phase(rsvdb_ctx* c, const double* A, int64_t K, int64_t M, int64_t lda, const double* Q, int64_t ldq, int N,
                         double* Z, int64_t ldz, int transpose_out, bool reduce, const Centering* cen = nullptr) {
  {
    PhaseTimer pt(c, PH_GEMM_AT);
    int k = 0;
    RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, A, K, M, lda, Q, ldq, N, Z, ldz, transpose_out, &k));
    c->launches += k;
  }
  if (cen) {                                                 // Ac^T Q = D (A^T Q - mu (1^T Q)); linear, so it is applied per shard
    if (transpose_out) return fail(c, -1, "implicit centring expects the M x N output layout");
    PhaseTimer po(c, PH_OTHER);
    double* sv = cen_xs(c, M) + PcaScratch::pad((size_t)M * N) + PcaScratch::pad((size_t)N);
    RSVDB_TRY(weighted_colsum(c, Q, ldq, K, N, nullptr, sv));
    RSVDB_TRY(rank1_correct(c, Z, ldz, M, N, cen->mu, sv, cen->inv_sd));
  }
  if (reduce && c->nranks > 1) {
    // A^T Q = sum over row shards of A_p^T Q_p; Z is contiguous (ldz == M or N) by construction in this file
    PhaseTimer pc(c, PH_COMM);
    RSVDB_TRY(comm_allreduce_sum(c, Z, transpose_out ? (size_t)ldz * M : (size_t)ldz * N));   // padding rows ride along
  }
  return 0;
}

// First pass with A arriving over PCIe: row block i is copied on the side stream while block i-1 is multiplied.
static int upload_and_first_pass(rsvdb_ctx* c, double* A, int64_t m, int64_t n, int64_t lda, const HostUpload& up,
                                 const double* Omega, int64_t ldo, int l, double* Q, int64_t ldq) {
  constexpr int kMaxBlocks = 16;                       // side_ev[0..15]
  const size_t bytes = (size_t)m * n * sizeof(double);
  int nb = (int)std::min<size_t>(kMaxBlocks, bytes >> 26);   // blocks of >= 64 MB, else the split costs more than it hides
  if (nb < 2 || m < 512) nb = 1;
  int64_t rows_b = ((m + nb - 1) / nb + 255) & ~int64_t(255);
  RSVDB_CUDA(c, cudaEventRecord(c->side_ev[16], c->stream));           // the device buffer may still be read by earlier work
  RSVDB_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->side_ev[16], 0));
  int b = 0;
  for (int64_t r0 = 0; r0 < m; r0 += rows_b, ++b) {
    const int64_t r = std::min(rows_b, m - r0);
    RSVDB_CUDA(c, upload_block(c, c->side_stream, A + r0, lda, up.A + r0, up.lda, r, n));   // pageable sources: pinned staging ring
    RSVDB_CUDA(c, cudaEventRecord(c->side_ev[b], c->side_stream));
    RSVDB_CUDA(c, cudaStreamWaitEvent(c->stream, c->side_ev[b], 0));
    RSVDB_TRY(gemm_an_phase(c, A + r0, r, n, lda, Omega, ldo, l, Q + r0, ldq));
  }
  return 0;
}

int range_finder(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                 int l, int q, double* Q, int64_t ldq, const HostUpload* up, const Centering* cen) {
  if (l <= 0 || q < 0) return fail(c, -1, "range_finder: l must be positive and q non-negative");
  c->chol_failed = false;                          // a new factorisation: the fast orthonormalisation gets its chance again
  // Z (n x l, even leading dimension: TMA needs 16-byte column strides) lives in tmp_ws at offset 0
  const int64_t ldz = even_ld(n);
  RSVDB_CUDA(c, c->tmp_ws.reserve(std::max<size_t>(c->tmp_ws.bytes, (size_t)ldz * l * sizeof(double))));
  double* Z = c->tmp_ws.ptr;
  if (cen && c->pca_ws.bytes < PcaScratch::total(n, l) * sizeof(double))
    return fail(c, -1, "range_finder: reserve pca_ws (PcaScratch::total) before building a Centering");
  if (up && up->A && m > 0) {
    if (cen) return fail(c, -1, "range_finder: the blocked upload and implicit centring are separate entry points");
    RSVDB_TRY(upload_and_first_pass(c, const_cast<double*>(A), m, n, lda, *up, Omega, ldo, l, Q, ldq));
  } else {
    RSVDB_TRY(gemm_an_phase(c, A, m, n, lda, Omega, ldo, l, Q, ldq, cen));     // Y = A * Omega          src/rSVD.cpp:59
  }
  RSVDB_TRY(orthonormalize(c, Q, m, l, ldq, true, nullptr));                   // Q = qr(Y).Q            :60-61
  for (int it = 0; it < q; ++it) {                                             // :62
    RSVDB_TRY(gemm_at_phase(c, A, m, n, lda, Q, ldq, l, Z, ldz, 0, true, cen));// Y = A^T * Q            :63
    RSVDB_TRY(orthonormalize(c, Z, n, l, ldz, false, nullptr));                // Q = qr(Y).Q  (n x l)   :64-65
    RSVDB_TRY(gemm_an_phase(c, A, m, n, lda, Z, ldz, l, Q, ldq, cen));         // Y = A * Q              :66
    RSVDB_TRY(orthonormalize(c, Q, m, l, ldq, true, nullptr));                 // Q = qr(Y).Q            :67-68
  }
  return 0;
}

int small_svd_jacobi(rsvdb_ctx* c, const double* M, int64_t ldm, const double* Mt, int64_t ldmt, int64_t r, int64_t cd,
                     double* U, int64_t ldu, double* S, double* V, int64_t ldv) {
  PhaseTimer pt(c, PH_SMALL_SVD);
  if (!c->in_rsvd) c->chol_failed = false;         // a stand-alone SVD<Jacobi>: its own factorisation
  const int64_t k = std::min(r, cd);
  if (k <= 0) return 0;
  if (k > 32768) return fail(c, -6, "SVD<Jacobi>: min(rows, cols) > 32768 is not supported (two k x k work matrices must fit in device memory)");
  int nl = 0;
  // scratch: tall copy (max(r,cd) x k), Uw (k x k), Zw (k x k), info
  const int64_t tall = std::max(r, cd);
  const size_t need = ((size_t)tall * k + 2 * (size_t)k * k + 16) * sizeof(double);
  RSVDB_CUDA(c, c->svd_ws.reserve(need));
  double* T = c->svd_ws.ptr; double* Uw = T + (size_t)tall * k; double* Zw = Uw + (size_t)k * k;
  int* info = reinterpret_cast<int*>(Zw + (size_t)k * k);
  GemmWorkspace& jws = c->qr2_ws;   // global-memory Jacobi scratch for k > ~116 (qr2_ws is idle here)
  if (r == cd) {
    // no preconditioner (include/SVD_class.hpp:110-123 takes neither branch)
    const double* W = M ? M : Mt; const int64_t ldw = M ? ldm : ldmt;
    RSVDB_CUDA(c, jacobi_svd_square(jws, c->stream, W, ldw, (int)k, M ? 0 : 1, U, ldu, S, V, ldv, info, &nl));
  } else if (r > cd) {
    // QR(M) -> work = R (c x c), U = Q_M * Uw, V = Zw                     (:110-115)
    if (M) { RSVDB_TRY(copy2d(c, M, ldm, T, r, r, (int)cd)); }
    else { RSVDB_TRY(transpose2d(c, Mt, ldmt, T, r, cd, r)); }
    const double* R = nullptr;
    RSVDB_TRY(orthonormalize(c, T, r, (int)cd, r, false, &R));
    RSVDB_CUDA(c, jacobi_svd_square(jws, c->stream, R, cd, (int)k, 0, Uw, k, S, V, ldv, info, &nl));
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, T, r, k, r, Uw, k, (int)k, U, ldu, &nl));
  } else {
    // QR(M^T) -> work = R^T (r x r), U = Uw, V = Q_{M^T} * Zw              (:116-123)
    if (Mt) { RSVDB_TRY(copy2d(c, Mt, ldmt, T, cd, cd, (int)r)); }
    else { RSVDB_TRY(transpose2d(c, M, ldm, T, cd, r, cd)); }
    const double* R = nullptr;
    RSVDB_TRY(orthonormalize(c, T, cd, (int)r, cd, false, &R));
    RSVDB_CUDA(c, jacobi_svd_square(jws, c->stream, R, r, (int)k, 1, U, ldu, S, Zw, k, info, &nl));
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, T, cd, k, cd, Zw, k, (int)k, V, ldv, &nl));
  }
  c->launches += nl;
  c->d_svd_info = info;   // sweeps / rotations are fetched lazily by rsvdb_last_svd_info()
  return 0;
}

int rsvd_device(rsvdb_ctx* c, const double* A, int64_t m, int64_t n, int64_t lda, const double* Omega, int64_t ldo,
                int l, int q, int method, double* U, int64_t ldu, double* S, double* V, int64_t ldv, uint64_t seed,
                const HostUpload* up, const Centering* cen) {
  if (method != 0 && method != 1 && method != 2) return fail(c, -1, "Unsupported SVD method");   // src/rSVD.cpp:122-123
  if (l <= 0 || n <= 0 || m < 0) return fail(c, -1, "rSVD: bad shape");
  const int64_t k = std::min<int64_t>(l, n);
  struct InRsvd { rsvdb_ctx* c; InRsvd(rsvdb_ctx* x) : c(x) { c->in_rsvd = true; } ~InRsvd() { c->in_rsvd = false; } } in_rsvd(c);
  // tmp_ws layout: [Z / Bt : n x l][Q : m x l][Ut : l x k]; even leading dimensions and even offsets keep every
  // sub-buffer TMA-addressable (16-byte base and column stride) whatever the parity of m, n and l
  const int64_t ldz = even_ld(n), ldq = even_ld(m), ldut = even_ld(l);
  const size_t need = ((size_t)ldz * l + (size_t)ldq * l + (size_t)ldut * l + 64) * sizeof(double);
  RSVDB_CUDA(c, c->tmp_ws.reserve(need));
  double* Bt = c->tmp_ws.ptr;
  double* Q = Bt + (size_t)ldz * l;
  double* Ut = Q + (size_t)ldq * l;
  RSVDB_TRY(range_finder(c, A, m, n, lda, Omega, ldo, l, q, Q, ldq, up, cen));      // Stage A          src/rSVD.cpp:84-85
  RSVDB_TRY(gemm_at_phase(c, A, m, n, lda, Q, ldq, l, Bt, ldz, 0, true, cen));      // B^T = A^T Q      :89 (stored transposed)
  if (method == 1) {
    RSVDB_TRY(small_svd_power_t(c, Bt, ldz, l, n, 0, seed, Ut, ldut, l, S, V, ldv, nullptr));   // SVD<Power>(B)    :105-112
  } else {
    RSVDB_TRY(small_svd_jacobi(c, nullptr, 0, Bt, ldz, l, n, Ut, ldut, S, V, ldv));  // SVD<method>(B)   :96-121
  }
  {
    PhaseTimer pt(c, PH_OTHER);                                                      // U = Q * Utilde   :128
    int nl = 0;
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, Q, m, l, ldq, Ut, ldut, (int)k, U, ldu, &nl));
    c->launches += nl;
  }
  return 0;
}

}  // namespace rsvdb
