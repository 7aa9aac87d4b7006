// Guarded CholeskyQR2: the fast way to orthonormalise a tall sketch whose columns are numerically independent.
//
// The reference orthonormalises every sketch with Eigen::HouseholderQR and forms the thin Q (src/rSVD.cpp:60-61, :64-65,
// :67-68; the QR preconditioner of SVD<Jacobi>, include/SVD_class.hpp:110-123).  Only the orthonormal basis (and, for the
// preconditioner, an upper-triangular R with Y = Q R) enters what follows.  A Householder TSQR pays a chain of l dependent
// reflector steps per tree level (tsqr.cu: ~1.3-3 us each, 2.9 ms for a 200000 x 100 panel); the same basis comes out of
// two rounds of
//     G = Y^T Y   (DMMA, split over all SMs; one l x l all-reduce when Y is row-sharded)
//     G = R^T R,  X = R^-1   (one CTA, k_chol_inv below)
//     Y <- Y X    (DMMA)
// in a quarter of the time -- PROVIDED Y is not too ill-conditioned: the first round leaves ||Q1^T Q1 - I|| ~ kappa(Y)^2 u,
// and the second round restores orthogonality to O(u) only if that is well below 1 (Yamamoto et al. 2015).  The guard is
// measured, not estimated: the second Gram matrix IS Q1^T Q1, its distance from I is computed while it is loaded, and the
// host reads it (one 32-byte read-back per QR) before Y is overwritten.  When that distance is below 1e-9 -- the usual case --
// and no R is wanted, round 2 needs no Cholesky at all: X = I - E/2 is (I + E)^(-1/2) to 4e-19.  A Cholesky breakdown (non-positive pivot: rank-
// deficient sketches such as the reference's configs 1 and 4) or a distance above CHOL_DEV_TOL leaves Y untouched and the
// caller runs the Householder TSQR exactly as before.  Nothing is ever computed on the CPU.
#include "pipeline.cuh"

#include <algorithm>
#include <cstdlib>

#include "comm.cuh"

namespace rsvdb {

namespace {

constexpr int CH_THREADS = 1024;
constexpr int CHOL_MIN_L = 16;
constexpr int CHOL_MAX_L = 128;            // 32 x 32 threads own up to 4 x 4 entries of the l x l work matrix each
constexpr double CHOL_FIRST_ORDER_TOL = 1e-9;  // ||Q1^T Q1 - I||_F below which round 2 uses X = I - E/2 instead of a Cholesky
constexpr double CHOL_DEV_TOL = 0.05;      // ||Q1^T Q1 - I||_F accepted before the second round (theory: <= 5/64)

// G (l x l, symmetric, column-major ldg) = L L^T.  Right-looking; the same eliminations run on an identity, so L^-1 is
// finished together with L:
//   step j:  s = 1 / sqrt(G[j][j]);   column j of L = G[:, j] s;   row j of L^-1 = W[j, :] s;
//            for i > j:  G[i, c] -= L[i][j] L[c][j]  (j < c <= i);   W[i, c] -= L[i][j] Linv[j][c]  (c <= j)
// G (columns > j) and W (columns <= j; its unit diagonal is implicit) share one lower-triangular work matrix that never
// leaves the REGISTERS: thread (ty, tx) owns the entries (i, c) with i = ty mod 32, c = tx mod 32 (every warp keeps a row in
// every 32-row block, so the per-step work of a warp stays at <= NT rows to the end -- the step time is the longest warp's
// instruction stream, measured: blocking the rows per warp instead was 40 % slower).  A step needs one vector from other
// threads -- v[c] = W[j][c] (c < j), 1 (c = j), G[c][j] (c > j) -- which its owners publish to shared memory (double-buffered:
// one block barrier per step); then every entry is m(i, c) += v[i] (-s^2 v[c]), the owners of column j having zeroed theirs.
// The loop over j is unrolled by 32-column blocks so that all register indices are compile-time and finished blocks drop out.
// Outputs: X = L^-T = R^-1 (upper triangular, ldx), R = L^T (upper triangular, ldr, optional).
// info[0] = 1 + index of the first non-positive pivot (0: none);  info[1] = ||G - I||_F^2.
template <int NT>
__global__ void __launch_bounds__(CH_THREADS, 1)
k_chol_inv(const double* __restrict__ G, int ldg, int l, double* __restrict__ X, int ldx, double* __restrict__ R, int ldr,
           double* __restrict__ info, int first_order_ok) {
  __shared__ double vec[2][32 * NT];
  __shared__ double piv[32 * NT], sinv[32 * NT];          // pivots and their rsqrt (computed once, by the owner)
  __shared__ double s_red[32];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  double m[NT][NT];
  double dev = 0.0;
#pragma unroll
  for (int a = 0; a < NT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) {
      const int i = ty + 32 * a, c = tx + 32 * b;
      double g = 0.0;
      if (i < l && c <= i) {
        g = G[(size_t)i * ldg + c];                          // (c, i) of the symmetric matrix: coalesced over tx
        const double d = g - ((i == c) ? 1.0 : 0.0);
        dev = fma((i == c) ? d : 2.0 * d, d, dev);
        if (i > c) { X[(size_t)c * ldx + i] = 0.0; if (R) R[(size_t)c * ldr + i] = 0.0; }
      }
      m[a][b] = g;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dev += __shfl_xor_sync(0xffffffffu, dev, o);
  if (tx == 0) s_red[ty] = dev;
  __syncthreads();
  if (first_order_ok) {
    // Second round on a sketch that is already orthonormal to ||E||_F <= 1e-9 (G = I + E; the usual case: kappa(Y)^2 u is far
    // below that): X = I - E/2 is G^(-1/2) up to 3/8 E^2 <= 4e-19, so the l sequential pivots are skipped.  X is symmetric
    // instead of triangular, which is why the caller allows this only when it does not ask for R.
    double t = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) t += s_red[w];          // same order on every thread: uniform decision
    if (t <= CHOL_FIRST_ORDER_TOL * CHOL_FIRST_ORDER_TOL) {
#pragma unroll
      for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          const int i = ty + 32 * a, c = tx + 32 * b;
          if (i < l && c <= i) {
            const double x = (i == c) ? 1.0 - 0.5 * (m[a][b] - 1.0) : -0.5 * m[a][b];
            X[(size_t)i * ldx + c] = x;
            if (i != c) X[(size_t)c * ldx + i] = x;
          }
        }
      if (tid == 0) { info[1] = t; info[0] = 0.0; }
      return;
    }
  }

  int bad = 0;
#pragma unroll
  for (int bj = 0; bj < NT; ++bj) {
    for (int tj = 0; tj < 32; ++tj) {
      const int j = 32 * bj + tj;
      if (j >= l) break;
      const int buf = j & 1;
      // ---- publish v for step j: column j below the diagonal (lane tj of every warp), the pivot, row j left of it (warp tj)
      if (tx == tj) {
        if (ty > tj) { vec[buf][ty + 32 * bj] = m[bj][bj]; m[bj][bj] = 0.0; }
#pragma unroll
        for (int a = bj + 1; a < NT; ++a) { vec[buf][ty + 32 * a] = m[a][bj]; m[a][bj] = 0.0; }   // rows >= l: never read
      }
      if (ty == tj) {
        if (tx == tj) { const double pv = m[bj][bj]; vec[buf][j] = 1.0; piv[j] = pv; sinv[j] = (pv > 0.0) ? rsqrt(pv) : 0.0; }
#pragma unroll
        for (int b = 0; b <= bj; ++b) { const int c = tx + 32 * b; if (c < j) vec[buf][c] = m[bj][b]; }
      }
      __syncthreads();
      // ---- step j
      const double d = piv[j];
      if (!(d > 0.0) || !(d < 1.7e308)) { bad = j + 1; goto finished; }            // uniform: every thread reads the same word
      const double s = sinv[j];
      if (R) { if (tid >= j && tid < l) R[(size_t)tid * ldr + j] = (tid == j ? d : vec[buf][tid]) * s; }   // R(j, i) = L(i, j)
      const double ns2 = -s * s;
      double lc[NT];
#pragma unroll
      for (int b = 0; b < NT; ++b) lc[b] = vec[buf][tx + 32 * b] * ns2;             // columns >= l: junk into junk entries
      if (ty > tj) {
        const double vi = vec[buf][ty + 32 * bj];
#pragma unroll
        for (int b = 0; b <= bj; ++b) m[bj][b] = fma(vi, lc[b], m[bj][b]);
      }
#pragma unroll
      for (int a = bj + 1; a < NT; ++a) {
        const double vi = vec[buf][ty + 32 * a];                                    // rows >= l: junk rows, never published
#pragma unroll
        for (int b = 0; b <= a; ++b) m[a][b] = fma(vi, lc[b], m[a][b]);
      }
    }
  }
finished:
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) t += s_red[w];
    info[1] = t;
    info[0] = (double)bad;
  }
  if (bad) return;
  // X(c, i) = Linv(i, c) = W[i][c] s_i (c < i), s_i (c = i)
#pragma unroll
  for (int a = 0; a < NT; ++a) {
    const int i = ty + 32 * a;
    if (i < l) {
      const double si = sinv[i];
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        const int c = tx + 32 * b;
        if (c <= i) X[(size_t)i * ldx + c] = (c == i) ? si : m[a][b] * si;
      }
    }
  }
}

// R (l x l, upper triangular) = R2 * R1, both upper triangular
__global__ void k_tri_mul(const double* __restrict__ R2, const double* __restrict__ R1, int ld, int l, double* __restrict__ R, int ldr) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= l * l) return;
  const int i = e % l, j = e / l;
  double a = 0.0;
  for (int k = i; k <= j; ++k) a = fma(R2[(size_t)k * ld + i], R1[(size_t)j * ld + k], a);
  R[(size_t)j * ldr + i] = a;
}

int policy_default() {
  static int p = -1;
  if (p < 0) { const char* e = getenv("RSVDB_CHOLQR"); p = (e && atoi(e) == 0) ? 1 : 0; }
  return p;
}

}  // namespace

// Returns 0 and sets *done = true when Y holds the orthonormal factor; *done = false: Y is untouched, use the TSQR.
static int cholqr2_try(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool dist, const double** Rout, bool* done) {
  *done = false;
  cudaStream_t st = c->stream;
  const int64_t ldt = even_ld(rows);
  const int ldl = (int)even_ld(l);
  const size_t sq = (size_t)ldl * l;
  // [T rows x l][G][X1][X2][R1][R2][R][info 8]
  RSVDB_CUDA(c, c->chol_ws.reserve(((size_t)ldt * l + 6 * sq + 8) * sizeof(double)));
  double* T = c->chol_ws.ptr;
  double* G = T + (size_t)ldt * l; double* X1 = G + sq; double* X2 = X1 + sq;
  double* R1 = X2 + sq; double* R2 = R1 + sq; double* R = R2 + sq; double* info = R + sq;
  if (!c->chol_host) RSVDB_CUDA(c, cudaMallocHost(&c->chol_host, 8 * sizeof(double)));
  auto chol = [&](const double* Gm, double* Xk, double* Rk, double* inf, int fo) {
    if (l <= 32) k_chol_inv<1><<<1, CH_THREADS, 0, st>>>(Gm, ldl, l, Xk, ldl, Rk, ldl, inf, fo);
    else if (l <= 64) k_chol_inv<2><<<1, CH_THREADS, 0, st>>>(Gm, ldl, l, Xk, ldl, Rk, ldl, inf, fo);
    else if (l <= 96) k_chol_inv<3><<<1, CH_THREADS, 0, st>>>(Gm, ldl, l, Xk, ldl, Rk, ldl, inf, fo);
    else k_chol_inv<4><<<1, CH_THREADS, 0, st>>>(Gm, ldl, l, Xk, ldl, Rk, ldl, inf, fo);
  };
  int nl = 0;
  auto round = [&](const double* src, int64_t lds, double* dst, int64_t ldd, double* Xk, double* Rk, double* inf) -> int {
    RSVDB_CUDA(c, gemm_at(c->gemm_ws, st, c->nsm, src, rows, l, lds, src, lds, l, G, ldl, 0, &nl));       // G = src^T src
    if (dist) { PhaseTimer pc(c, PH_COMM); RSVDB_TRY(comm_allreduce_sum(c, G, sq)); }
    chol(G, Xk, Rk, inf, 0); ++nl;
    RSVDB_CUDA(c, cudaGetLastError());
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, st, c->nsm, src, rows, l, lds, Xk, ldl, l, dst, ldd, &nl));          // dst = src X
    return 0;
  };
  // round 1 into T, round 2's Gram and Cholesky, then the verdict -- Y is still intact at that point
  RSVDB_TRY(round(Y, ldy, T, ldt, X1, Rout ? R1 : nullptr, info));
  RSVDB_CUDA(c, gemm_at(c->gemm_ws, st, c->nsm, T, rows, l, ldt, T, ldt, l, G, ldl, 0, &nl));
  if (dist) { PhaseTimer pc(c, PH_COMM); RSVDB_TRY(comm_allreduce_sum(c, G, sq)); }
  chol(G, X2, Rout ? R2 : nullptr, info + 2, Rout ? 0 : 1); ++nl;                     // no R wanted: X2 may be the symmetric I - E/2
  RSVDB_CUDA(c, cudaGetLastError());
  RSVDB_CUDA(c, cudaMemcpyAsync(c->chol_host, info, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  RSVDB_CUDA(c, cudaStreamSynchronize(st));
  c->launches += nl; nl = 0;
  const double* h = c->chol_host;
  const bool ok = h[0] == 0.0 && h[2] == 0.0 && h[3] <= CHOL_DEV_TOL * CHOL_DEV_TOL;     // NaN compares false
  if (!ok) return 0;
  RSVDB_CUDA(c, gemm_an(c->gemm_ws, st, c->nsm, T, rows, l, ldt, X2, ldl, l, Y, ldy, &nl));                // Q = Q1 X2
  if (Rout) {
    k_tri_mul<<<(l * l + 255) / 256, 256, 0, st>>>(R2, R1, ldl, l, R, l); ++nl;                            // R = R2 R1, ld = l like the TSQR's
    RSVDB_CUDA(c, cudaGetLastError());
    *Rout = R;
  }
  c->launches += nl;
  *done = true;
  return 0;
}

int orthonormalize(rsvdb_ctx* c, double* Y, int64_t rows, int l, int64_t ldy, bool sharded, const double** R) {
  const bool dist = sharded && c->nranks > 1;
  const int policy = c->qr_policy < 0 ? policy_default() : c->qr_policy;
  // every condition below is identical on all ranks of a sharded panel (shard heights are not consulted when sharded)
  // (narrow sketches and single-leaf panels stay on the TSQR: its reflector chain is short there and CholeskyQR2 has ~0.1 ms of
  // fixed cost; measured break-even in tools/orth_check.py)
  const bool eligible = policy == 0 && !c->chol_failed && l >= CHOL_MIN_L && l <= CHOL_MAX_L && (dist || rows >= std::max<int64_t>(l, 512));
  if (eligible) {
    bool done = false;
    {
      PhaseTimer pt(c, PH_QR);
      RSVDB_TRY(cholqr2_try(c, Y, rows, l, ldy, dist, R, &done));
    }
    if (done) { ++c->qr_fast; return 0; }
    c->chol_failed = true;              // ill-conditioned or rank-deficient sketch: the rest of this factorisation stays on Householder
  }
  if (l > CHOL_MAX_L && policy == 0 && !c->wide_fast) {
    // wider than the Cholesky kernel: block Gram-Schmidt over <= 104-column blocks (qr_wide), each block through this function
    struct Scope { rsvdb_ctx* c; Scope(rsvdb_ctx* x) : c(x) { c->wide_fast = true; } ~Scope() { c->wide_fast = false; } } scope(c);
    return qr_inplace(c, Y, rows, l, ldy, sharded, R);
  }
  ++c->qr_householder;
  return qr_inplace(c, Y, rows, l, ldy, sharded, R);
}

}  // namespace rsvdb
