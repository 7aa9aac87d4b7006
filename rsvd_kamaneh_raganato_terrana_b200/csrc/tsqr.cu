// Communication-avoiding Householder QR (TSQR) of a tall-skinny column-major FP64 matrix, thin Q formed explicitly.
//
// Replaces `Eigen::HouseholderQR<MatrixXd> qr(Y); Q = qr.householderQ() * Identity(rows, l)` at reference
// src/rSVD.cpp:60-61,64-65,67-68 and the QR preconditioner of the small SVD at include/SVD_class.hpp:112-122.
// Same reflector convention as Eigen / LAPACK dlarfg: for x = [x0; t], beta = -sign(x0)*||x|| (sign(0) = +1),
// v = [1; t/(x0-beta)], tau = (beta-x0)/beta, and tau = 0 (no reflection) when ||t||^2 <= DBL_MIN -- so rank-deficient
// panels (reference config 1: a rank-2 matrix sketched with l = 16) behave like the reference's Householder QR; no
// Gram matrix is ever formed.
//
// Layout: the rows are cut into leaves of BR rows.  One CTA factors one leaf entirely in shared memory (column-major,
// BR x l), leaving the reflectors in place of Y and its l x l R in a stack; the stack ((#leaves * l) x l) is factored
// by the same kernel, recursively, until one leaf remains.  Q is then formed top-down: the CTA of leaf b applies its
// reflectors to [C_b; 0], where C_b is the b-th l x l block of the explicit Q of the level above (identity at the
// top, or -- multi-GPU -- this rank's block of the Q of the all-gathered R factors).
#include "dev_once.cuh"
#include "tsqr.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <vector>

namespace rsvdb {

namespace {

constexpr int QR_THREADS = 512;
constexpr int QR_WARPS = QR_THREADS / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Build the reflector of column `col` (diagonal row d = col) from the lane-distributed column values y (row i = lane + 32*ii),
// write the scaled essential part, beta on the diagonal and tau.  Executed by one full warp.
template <int RPL>
__device__ __forceinline__ void make_reflector(double* colp, const double (&y)[RPL], int d, int lane, double* tau_out) {
  double tail = 0.0, x0 = 0.0;
#pragma unroll
  for (int ii = 0; ii < RPL; ++ii) {
    const int i = lane + 32 * ii;
    if (i > d) tail += y[ii] * y[ii];
    if (i == d) x0 = y[ii];
  }
  tail = warp_sum(tail);
  x0 = warp_sum(x0);
  double beta, tau, scale;
  if (tail <= DBL_MIN) { tau = 0.0; beta = x0; scale = 0.0; }
  else {
    beta = sqrt(x0 * x0 + tail);
    if (x0 >= 0.0) beta = -beta;
    scale = 1.0 / (x0 - beta);
    tau = (beta - x0) / beta;
  }
#pragma unroll
  for (int ii = 0; ii < RPL; ++ii) {
    const int i = lane + 32 * ii;
    if (i > d) colp[i] = y[ii] * scale;
    else if (i == d) colp[i] = beta;
  }
  if (lane == 0) *tau_out = tau;
}

// Factor one leaf per CTA.  Y is rows x l (ldy); leaf b covers rows [b*BR, min(rows,(b+1)*BR)), zero-padded to BR.
// On exit the reflectors (essential parts, below the diagonal) overwrite Y, tau[b*l + j] holds the scalars and the
// l x l upper-triangular R of the leaf is stored at rows [b*l, (b+1)*l) of Rstack (ldr).
template <int BR>
__global__ void __launch_bounds__(QR_THREADS, 1)
k_house_factor(double* __restrict__ Y, long long ldy, long long rows, int l, double* __restrict__ tau_g,
               double* __restrict__ Rstack, long long ldr) {
  constexpr int RPL = BR / 32;
  extern __shared__ double sm[];
  double* S = sm;                 // l columns of BR
  double* tau_s = sm + (size_t)l * BR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * BR;
  const int nrows = (int)min((long long)BR, rows - r0);

  for (int k = warp; k < l; k += QR_WARPS) {
    const double* src = Y + (size_t)k * ldy + r0;
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      S[(size_t)k * BR + i] = (i < nrows) ? src[i] : 0.0;
    }
  }
  __syncthreads();
  const int nref = min(l, BR);
  if (warp == 0) {
    double y[RPL];
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) y[ii] = S[lane + 32 * ii];
    make_reflector<RPL>(S, y, 0, lane, &tau_s[0]);
  }
  __syncthreads();

  for (int j = 0; j < nref; ++j) {
    const double tau = tau_s[j];
    const int ii0 = j >> 5;                    // rows below 32*ii0 are above the diagonal for every lane
    double v[RPL];
    const double* vj = S + (size_t)j * BR;
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      v[ii] = (ii >= ii0) ? ((i > j) ? vj[i] : (i == j ? 1.0 : 0.0)) : 0.0;
    }
    for (int k = j + 1 + warp; k < l; k += QR_WARPS) {
      double* ck = S + (size_t)k * BR;
      double y[RPL];
      double dot = 0.0;
#pragma unroll
      for (int ii = 0; ii < RPL; ++ii) {
        y[ii] = (ii >= ii0) ? ck[lane + 32 * ii] : 0.0;
        dot = fma(v[ii], y[ii], dot);
      }
      const bool next = (k == j + 1) && (j + 1 < nref);
      if (tau != 0.0) {
        const double w = tau * warp_sum(dot);
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
          y[ii] = fma(-w, v[ii], y[ii]);
          if (ii >= ii0 && !next) ck[lane + 32 * ii] = y[ii];
        }
      }
      if (next) {
        // rows <= j of column j+1 are final R entries: they must be stored un-scaled
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
          const int i = lane + 32 * ii;
          if (ii >= ii0 && i <= j) ck[i] = y[ii];
        }
        double yy[RPL];
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) yy[ii] = (ii >= ii0) ? y[ii] : 0.0;
        // make_reflector only touches rows >= j+1; rows < 32*ii0 are untouched because y is 0 there and i <= d
        make_reflector<RPL>(ck, yy, j + 1, lane, &tau_s[j + 1]);
      }
    }
    __syncthreads();
  }

  // write back: reflectors below the diagonal -> Y, upper triangle -> Rstack, tau
  double* Rb = Rstack + (size_t)blockIdx.x * l;
  for (int k = warp; k < l; k += QR_WARPS) {
    double* dst = Y + (size_t)k * ldy + r0;
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      const double val = S[(size_t)k * BR + i];
      if (i < nrows) dst[i] = val;
      if (i < l) Rb[(size_t)k * ldr + i] = (i <= k && i < nref) ? val : 0.0;
    }
    if (BR < l) for (int i = BR + lane; i < l; i += 32) Rb[(size_t)k * ldr + i] = 0.0;
  }
  for (int j = threadIdx.x; j < l; j += QR_THREADS) tau_g[(size_t)blockIdx.x * l + j] = (j < nref) ? tau_s[j] : 0.0;
}

// Form the leaf's rows of the explicit thin Q:  Q_b = H_0 H_1 ... H_{nref-1} [C_b; 0].
// V (ldv) holds the leaf reflectors; C_b = rows [b*l, (b+1)*l) of Ctop (ldc), or the identity when Ctop == nullptr.
// Q may alias V: each CTA reads only its own rows of V and writes them after the last read.
template <int BR>
__global__ void __launch_bounds__(QR_THREADS, 1)
k_house_apply(const double* V, long long ldv, long long rows, int l, const double* __restrict__ tau_g,
              const double* __restrict__ Ctop, long long ldc, double* Q, long long ldq) {
  constexpr int RPL = BR / 32;
  extern __shared__ double sm[];
  double* S = sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * BR;
  const int nrows = (int)min((long long)BR, rows - r0);
  const int nref = min(l, BR);

  for (int k = warp; k < l; k += QR_WARPS) {
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      double c = 0.0;
      if (i < l) c = Ctop ? Ctop[(size_t)k * ldc + (size_t)blockIdx.x * l + i] : (i == k ? 1.0 : 0.0);
      S[(size_t)k * BR + i] = c;
    }
  }
  const double* taub = tau_g + (size_t)blockIdx.x * l;
  // prefetch reflector nref-1
  double vn[RPL];
  {
    const int j = nref - 1;
    const double* vj = V + (size_t)j * ldv + r0;
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      vn[ii] = (i > j && i < nrows) ? vj[i] : (i == j ? 1.0 : 0.0);
    }
  }
  __syncthreads();
  for (int j = nref - 1; j >= 0; --j) {
    const double tau = taub[j];
    const int ii0 = j >> 5;
    double v[RPL];
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) v[ii] = vn[ii];
    if (j > 0) {
      const int jn = j - 1;
      const double* vj = V + (size_t)jn * ldv + r0;
#pragma unroll
      for (int ii = 0; ii < RPL; ++ii) {
        const int i = lane + 32 * ii;
        vn[ii] = (i > jn && i < nrows) ? vj[i] : (i == jn ? 1.0 : 0.0);
      }
    }
    if (tau != 0.0) {
      for (int k = warp; k < l; k += QR_WARPS) {
        double* ck = S + (size_t)k * BR;
        double y[RPL];
        double dot = 0.0;
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
          y[ii] = (ii >= ii0) ? ck[lane + 32 * ii] : 0.0;
          dot = fma(v[ii], y[ii], dot);
        }
        const double w = tau * warp_sum(dot);
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii)
          if (ii >= ii0) ck[lane + 32 * ii] = fma(-w, v[ii], y[ii]);
      }
    }
    __syncthreads();
  }
  for (int k = warp; k < l; k += QR_WARPS) {
    double* dst = Q + (size_t)k * ldq + r0;
#pragma unroll
    for (int ii = 0; ii < RPL; ++ii) {
      const int i = lane + 32 * ii;
      if (i < nrows) dst[i] = S[(size_t)k * BR + i];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Blocked (compact-WY) Householder leaves: panels of 8 reflectors, trailing updates on the FP64 tensor cores.
//
// The unblocked leaf above is latency-bound (one block barrier and one shuffle reduction chain per column and per
// reflector).  Here a panel of 8 columns is factored by ONE warp entirely in registers (no block barrier inside the
// panel), its 8x8 triangular factor T is formed (Q_p = I - V T V^T, LAPACK dlarft forward/columnwise), and the other
// columns are updated panel-at-a-time with three small GEMMs per 8-column group, all mma.sync.m8n8k4.f64 (SASS DMMA):
//   (a) W = V^T C          8 x 8, reduction over the block rows        (the MMA does the cross-lane reduction)
//   (b) X = op(T) W        8 x 8                                        (op = T^T when factoring, T when forming Q)
//   (c) C = C - V X        rows x 8
// Shared-memory layout: column-major with leading dimension BR + 4 (== 4 mod 16) and n-slot permutation
// PI = {0,1,2,3,5,4,7,6}, which makes every fragment load/store of (a) and (c) bank-conflict free.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BQ_THREADS = 512;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_ROW_WARPS = 8;                 // warps that own a 32-row slab of the panel (BR = 256)

__device__ __forceinline__ int perm8(int x) { return (x < 4) ? x : (x ^ 1); }   // {0,1,2,3,5,4,7,6}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// C-fragment (value of row g, slots 2t / 2t+1 held in c0 / c1) -> B-fragment value M[k][slot g] for this lane.
__device__ __forceinline__ double frag_c_to_b(double c0, double c1, int k, int g) {
  const int src = k * 4 + (g >> 1);
  const double x = __shfl_sync(0xffffffffu, c0, src), y = __shfl_sync(0xffffffffu, c1, src);
  return (g & 1) ? y : x;
}

// Apply the block reflector of one panel (clean reflectors Vp: 8 columns, ld LDS, zero above row0; T: 8x8 column-major)
// to the columns [col_begin, col_end) of S, rows [row0, BR).  transposeT: C <- (I - V T^T V^T) C, else (I - V T V^T) C.
template <int BR>
__device__ __forceinline__ void block_reflect(double* S, const double* Vp, const double* Tsm, int row0, int col_begin, int col_end,
                                              int warp, int nwarps, int lane, bool transposeT) {
  constexpr int LDS = BR + 4;
  const int g = lane >> 2, t = lane & 3;
  for (int n0 = col_begin + 8 * warp; n0 < col_end; n0 += 8 * nwarps) {
    // (a) W = Vp^T C  (four independent accumulator chains hide the DMMA latency)
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const int colB = n0 + perm8(g);
    const bool bval = colB < col_end;
    const double* cb = S + (size_t)(bval ? colB : col_begin) * LDS + row0 + t;
    const double* va = Vp + (size_t)g * LDS + row0 + t;
    const int nk = (BR - row0) >> 2;                                   // k4 steps (even)
    for (int k = 0; k < nk; k += 4) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool on = k + u < nk;
        av[u] = on ? va[4 * (k + u)] : 0.0;
        bv[u] = (on && bval) ? cb[4 * (k + u)] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(acc[u], av[u], bv[u]);
    }
    const double w0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);  // W[g][slot 2t]
    const double w1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);  // W[g][slot 2t+1]
    // (b) X = op(T) W
    const double bw0 = frag_c_to_b(w0, w1, t, g), bw1 = frag_c_to_b(w0, w1, t + 4, g);
    const double t0 = transposeT ? Tsm[t + 8 * g] : Tsm[g + 8 * t];
    const double t1 = transposeT ? Tsm[(t + 4) + 8 * g] : Tsm[g + 8 * (t + 4)];
    double x[2] = {0.0, 0.0};
    dmma884(x, t0, bw0);
    dmma884(x, t1, bw1);                                               // X[g][slot 2t], X[g][slot 2t+1]
    // (c) C -= Vp X
    const double bx0 = -frag_c_to_b(x[0], x[1], t, g), bx1 = -frag_c_to_b(x[0], x[1], t + 4, g);
    const int col0 = n0 + perm8(2 * t), col1 = n0 + perm8(2 * t + 1);
    const bool v0 = col0 < col_end, v1 = col1 < col_end;
    double* p0 = S + (size_t)(v0 ? col0 : col_begin) * LDS + row0 + g;
    double* p1 = S + (size_t)(v1 ? col1 : col_begin) * LDS + row0 + g;
    const double* a0p = Vp + (size_t)t * LDS + row0 + g;
    const double* a1p = Vp + (size_t)(t + 4) * LDS + row0 + g;
    const int nblk = (BR - row0) >> 3;                                 // 8-row blocks
    for (int b0 = 0; b0 < nblk; b0 += 4) {
      double c[4][2], x0[4], x1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool on = b0 + u < nblk; const int r = 8 * (b0 + u);
        c[u][0] = (on && v0) ? p0[r] : 0.0; c[u][1] = (on && v1) ? p1[r] : 0.0;
        x0[u] = on ? a0p[r] : 0.0; x1[u] = on ? a1p[r] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(c[u], x0[u], bx0);
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(c[u], x1[u], bx1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool on = b0 + u < nblk; const int r = 8 * (b0 + u);
        if (on && v0) p0[r] = c[u][0];
        if (on && v1) p1[r] = c[u][1];
      }
    }
  }
}

// Transposing warp reduction of 8 per-lane values: after three halving exchanges (xor 16, 8, 4) every lane owns ONE column,
// two more butterfly steps finish the sum.  9 shuffles instead of 40; returns the total of column `col` (valid in all lanes,
// lanes with (lane & 3) == 0 publish it).
__device__ __forceinline__ double reduce8_transposed(const double (&p)[8], int lane, int* col) {
  double q[4], r2[2], s;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = u16 ? p[i] : p[i + 4], keep = u16 ? p[i + 4] : p[i];
    q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = u8 ? q[i] : q[i + 2], keep = u8 ? q[i + 2] : q[i];
    r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double send = u4 ? r2[0] : r2[1], keep = u4 ? r2[1] : r2[0];
    s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  *col = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
  return s;
}

// ------------------------------------------------------------------------------------------------------------------
// Look-ahead factor kernel.  The 16 warps form two teams: the ROW team (warps 0-7, one 32-row slab each) updates the
// NEXT panel's 8 columns with the current panel's reflectors (cooperatively: per-slab partial V^T C on the tensor
// cores, summed through shared memory) and factors that panel right away, while the UPDATE team (warps 8-15) applies
// the current panel to the remaining columns.  The reflector chain of panel p+1 thus overlaps the trailing update of
// panel p; the teams meet at one block barrier per panel, the row team synchronises internally on named barrier 1.
// Reflectors are read in place from S (plus a clean 8 x 8 top block), double-buffered together with T.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// clean reflector value V[row][r] of the panel starting at column / row c0 (rows in [c0, c0+8) come from Vtop)
template <int BR>
__device__ __forceinline__ double clean_v(const double* S, const double* Vtop, int c0, int row, int r) {
  constexpr int LDS = BR + 4;
  if (row < c0) return 0.0;
  if (row < c0 + 8) return Vtop[r * 8 + (row - c0)];
  return S[(size_t)(c0 + r) * LDS + row];
}

// Row team: factor the panel [c0, c0+pb) (thread = one row).  Writes S (storage form), Vtop, Tsm, Tglob, tau_s.
template <int BR>
__device__ __forceinline__ void panel_factor_la(double* S, double* Vtop, double* Tsm, double* tau_s, double* Tglob, double* scratch,
                                                int c0, int pb, int warp, int lane) {
  constexpr int LDS = BR + 4;
  double* red = scratch;            // [2][8 values][8 warps]
  double* drow = scratch + 128;     // [2][8]
  double* Gs = scratch + 144;       // [8 warps][64]
  double* Gtot = scratch + 656;     // [64]
  const int i = 32 * warp + lane;
  double a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = (c < pb) ? S[(size_t)(c0 + c) * LDS + i] : 0.0;
  double tau[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    tau[r] = 0.0;
    if (r < pb) {
      const int d = c0 + r;
      const int b = r & 1;
      double p[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) p[c] = (c >= r && i > d) ? a[r] * a[c] : 0.0;
      {
        int colr; const double tcol = reduce8_transposed(p, lane, &colr);
        if ((lane & 3) == 0) red[b * 64 + colr * 8 + warp] = tcol;
      }
      if (i == d) {
#pragma unroll
        for (int c = 0; c < 8; ++c) drow[b * 8 + c] = a[c];
      }
      bar_rows();
      // cross-warp sum of the 8 x 8 partials: lane L reads ONE 16-byte pair (partials 2L, 2L+1 of column L >> 2), two butterfly
      // steps finish the column, then the totals are broadcast -- 1 LDS.128 + 10 shuffles instead of 32 LDS.128 + 56 adds per thread
      // (the read-back of all 64 partials by every thread was 35 % of the panel's ncu samples)
      double tot[8];
      {
        const double2 pr = reinterpret_cast<const double2*>(red + b * 64)[lane];
        double s2 = pr.x + pr.y;
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
#pragma unroll
        for (int c = 0; c < 8; ++c) tot[c] = (c >= r) ? __shfl_sync(0xffffffffu, s2, 4 * c) : 0.0;
      }
      const double tail = tot[r], x0 = drow[b * 8 + r];
      double beta, scale;
      if (tail <= DBL_MIN) { tau[r] = 0.0; beta = x0; scale = 0.0; }
      else {
        const double n2 = fma(x0, x0, tail);
        const double inrm = rsqrt(n2);
        const double nrm = n2 * inrm;
        const double ax = fabs(x0);
        beta = (x0 >= 0.0) ? -nrm : nrm;
        tau[r] = fma(ax, inrm, 1.0);
        const double rc = __drcp_rn(ax + nrm);
        scale = (x0 >= 0.0) ? rc : -rc;
      }
      const double v = (i > d) ? a[r] * scale : (i == d ? 1.0 : 0.0);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c > r && c < pb) {
          const double w = tau[r] * fma(scale, tot[c], drow[b * 8 + c]);
          a[c] = fma(-w, v, a[c]);
        }
      }
      if (i > d) a[r] = v; else if (i == d) a[r] = beta;
    }
  }
  // storage form back to S; clean top block (rows c0 .. c0+7) to Vtop
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c < pb) S[(size_t)(c0 + c) * LDS + i] = a[c];
    if (i >= c0 && i < c0 + 8) {
      const int d = c0 + c;
      Vtop[c * 8 + (i - c0)] = (c < pb) ? ((i > d) ? a[c] : (i == d ? 1.0 : 0.0)) : 0.0;
    }
  }
  __syncwarp();
  // Gram partial of this warp's 32 rows: G = V^T V on the tensor cores (each warp reads only rows it wrote itself)
  {
    const int g = lane >> 2, t = lane & 3;
    double acc[2] = {0.0, 0.0};
    if (32 * warp + 31 >= c0 && g < pb) {
      // (columns g >= pb do not exist: their clean reflector is zero)
    }
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
      const int row = 32 * warp + k0 + t;
      const double x = (g < pb) ? clean_v<BR>(S, Vtop, c0, row, g) : 0.0;
      dmma884(acc, x, x);
    }
    Gs[warp * 64 + g + 8 * (2 * t)] = acc[0];
    Gs[warp * 64 + g + 8 * (2 * t + 1)] = acc[1];
  }
  bar_rows();
  if (threadIdx.x < 64) {
    double s2 = 0.0;
#pragma unroll
    for (int w = 0; w < BQ_ROW_WARPS; ++w) s2 += Gs[w * 64 + threadIdx.x];
    Gtot[threadIdx.x] = s2;
  }
  bar_rows();
  if (threadIdx.x < 8) {
    const int srow = threadIdx.x;
    double trow[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      double val = 0.0;
      if (c == srow) val = tau[c];
      else if (c > srow) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) if (r < c && r >= srow) acc = fma(trow[r], Gtot[r + 8 * c], acc);
        val = -tau[c] * acc;
      }
      trow[c] = val;
      Tsm[srow + 8 * c] = val;
      Tglob[srow + 8 * c] = val;
    }
    double tv = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c == srow) tv = tau[c];
    if (srow < pb) tau_s[c0 + srow] = tv;
  }
}

// Row team: apply the reflectors of the panel at c0 to the 8 columns starting at n0 (the next panel), slab by slab.
template <int BR>
__device__ __forceinline__ void coop_update(double* S, const double* Vtop, const double* Tsm, double* Ws, int c0, int n0, int col_end,
                                            int warp, int lane) {
  constexpr int LDS = BR + 4;
  const int g = lane >> 2, t = lane & 3;
  const int rbase = 32 * warp;
  const bool any = rbase + 31 >= c0;
  const int colB = n0 + perm8(g);
  const bool bval = colB < col_end;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  if (any) {
    const double* cb = S + (size_t)(bval ? colB : n0) * LDS + rbase + t;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
      const double av = clean_v<BR>(S, Vtop, c0, rbase + k0 + t, g);
      const double bv = bval ? cb[k0] : 0.0;
      dmma884(acc[(k0 >> 2) & 1], av, bv);
    }
  }
  Ws[warp * 64 + g + 8 * (2 * t)] = acc[0][0] + acc[1][0];
  Ws[warp * 64 + g + 8 * (2 * t + 1)] = acc[0][1] + acc[1][1];
  bar_rows();
  if (!any) return;
  double w0 = 0.0, w1 = 0.0;
#pragma unroll
  for (int w = 0; w < BQ_ROW_WARPS; ++w) { w0 += Ws[w * 64 + g + 8 * (2 * t)]; w1 += Ws[w * 64 + g + 8 * (2 * t + 1)]; }
  const double bw0 = frag_c_to_b(w0, w1, t, g), bw1 = frag_c_to_b(w0, w1, t + 4, g);
  double x[2] = {0.0, 0.0};
  dmma884(x, Tsm[t + 8 * g], bw0);                 // X = T^T W
  dmma884(x, Tsm[(t + 4) + 8 * g], bw1);
  const double bx0 = -frag_c_to_b(x[0], x[1], t, g), bx1 = -frag_c_to_b(x[0], x[1], t + 4, g);
  const int col0 = n0 + perm8(2 * t), col1 = n0 + perm8(2 * t + 1);
  const bool v0 = col0 < col_end, v1 = col1 < col_end;
  double* p0 = S + (size_t)(v0 ? col0 : n0) * LDS + g;
  double* p1 = S + (size_t)(v1 ? col1 : n0) * LDS + g;
  double c[4][2], xa[4], xb[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r0 = rbase + 8 * u;
    const bool on = r0 + 7 >= c0;                  // c0 is a multiple of 8: the block is entirely below or entirely at/after c0
    c[u][0] = (on && v0) ? p0[r0] : 0.0; c[u][1] = (on && v1) ? p1[r0] : 0.0;
    xa[u] = on ? clean_v<BR>(S, Vtop, c0, r0 + g, t) : 0.0;
    xb[u] = on ? clean_v<BR>(S, Vtop, c0, r0 + g, t + 4) : 0.0;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) dmma884(c[u], xa[u], bx0);
#pragma unroll
  for (int u = 0; u < 4; ++u) dmma884(c[u], xb[u], bx1);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r0 = rbase + 8 * u;
    const bool on = r0 + 7 >= c0;
    if (on && v0) p0[r0] = c[u][0];
    if (on && v1) p1[r0] = c[u][1];
  }
}

// Update team: block_reflect with the reflectors read in place from S (+ Vtop for the panel's first 8 rows).
template <int BR>
__device__ __forceinline__ void block_reflect_s(double* S, const double* Vtop, const double* Tsm, int c0, int col_begin, int col_end,
                                                int warp, int nwarps, int lane) {
  constexpr int LDS = BR + 4;
  const int g = lane >> 2, t = lane & 3;
  for (int n0 = col_begin + 8 * warp; n0 < col_end; n0 += 8 * nwarps) {
    double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    const int colB = n0 + perm8(g);
    const bool bval = colB < col_end;
    const double* cb = S + (size_t)(bval ? colB : col_begin) * LDS + c0 + t;
    const double* va = S + (size_t)(c0 + g) * LDS + c0 + t;          // rows >= c0 + 8
    const double* vt = Vtop + g * 8 + t;                             // rows c0 .. c0+7
    const int nk = (BR - c0) >> 2;
    for (int k = 0; k < nk; k += 4) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kk = k + u; const bool on = kk < nk;
        av[u] = on ? (kk < 2 ? vt[4 * kk] : va[4 * kk]) : 0.0;
        bv[u] = (on && bval) ? cb[4 * kk] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(acc[u], av[u], bv[u]);
    }
    const double w0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
    const double w1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
    const double bw0 = frag_c_to_b(w0, w1, t, g), bw1 = frag_c_to_b(w0, w1, t + 4, g);
    double x[2] = {0.0, 0.0};
    dmma884(x, Tsm[t + 8 * g], bw0);
    dmma884(x, Tsm[(t + 4) + 8 * g], bw1);
    const double bx0 = -frag_c_to_b(x[0], x[1], t, g), bx1 = -frag_c_to_b(x[0], x[1], t + 4, g);
    const int col0 = n0 + perm8(2 * t), col1 = n0 + perm8(2 * t + 1);
    const bool v0 = col0 < col_end, v1 = col1 < col_end;
    double* p0 = S + (size_t)(v0 ? col0 : col_begin) * LDS + c0 + g;
    double* p1 = S + (size_t)(v1 ? col1 : col_begin) * LDS + c0 + g;
    const double* a0p = S + (size_t)(c0 + t) * LDS + c0 + g;
    const double* a1p = S + (size_t)(c0 + t + 4) * LDS + c0 + g;
    const int nblk = (BR - c0) >> 3;
    for (int b0 = 0; b0 < nblk; b0 += 4) {
      double c[4][2], x0[4], x1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int bb = b0 + u; const bool on = bb < nblk; const int r = 8 * bb;
        c[u][0] = (on && v0) ? p0[r] : 0.0; c[u][1] = (on && v1) ? p1[r] : 0.0;
        x0[u] = on ? (bb == 0 ? Vtop[t * 8 + g] : a0p[r]) : 0.0;
        x1[u] = on ? (bb == 0 ? Vtop[(t + 4) * 8 + g] : a1p[r]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(c[u], x0[u], bx0);
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884(c[u], x1[u], bx1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int bb = b0 + u; const bool on = bb < nblk; const int r = 8 * bb;
        if (on && v0) p0[r] = c[u][0];
        if (on && v1) p1[r] = c[u][1];
      }
    }
  }
}

template <int BR>
__global__ void __launch_bounds__(BQ_THREADS, 1)
k_house_factor_la(double* __restrict__ Y, long long ldy, long long rows, int l, double* __restrict__ tau_g,
                  double* __restrict__ Rstack, long long ldr, double* __restrict__ Tg) {
  constexpr int LDS = BR + 4;
  extern __shared__ double sm[];
  double* S = sm;
  double* Vtop = S + (size_t)l * LDS;              // [2][64]
  double* Tsm = Vtop + 128;                        // [2][64]
  double* scratch = Tsm + 128;                     // 720
  double* Ws = scratch + 720;                      // [8][64]
  double* tau_s = Ws + 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool roww = warp < BQ_ROW_WARPS;
  const long long r0 = (long long)blockIdx.x * BR;
  const int nrows = (int)min((long long)BR, rows - r0);
  const int npanels = (l + 7) / 8;

  for (int k = warp; k < l; k += BQ_WARPS) {
    const double* src = Y + (size_t)k * ldy + r0;
    for (int i = lane; i < BR; i += 32) S[(size_t)k * LDS + i] = (i < nrows) ? src[i] : 0.0;
  }
  __syncthreads();
  double* Tblock = Tg + (size_t)blockIdx.x * npanels * 64;
  if (roww) panel_factor_la<BR>(S, Vtop, Tsm, tau_s, Tblock, scratch, 0, min(8, l), warp, lane);
  __syncthreads();
  for (int p = 0; p < npanels; ++p) {
    const int c0 = 8 * p, pb = min(8, l - c0), buf = p & 1;
    const int n0 = c0 + pb;                        // first column of the next panel
    if (roww) {
      if (n0 < l) {
        coop_update<BR>(S, Vtop + buf * 64, Tsm + buf * 64, Ws, c0, n0, min(l, n0 + 8), warp, lane);
        __syncwarp();
        panel_factor_la<BR>(S, Vtop + (buf ^ 1) * 64, Tsm + (buf ^ 1) * 64, tau_s, Tblock + (size_t)(p + 1) * 64, scratch, n0, min(8, l - n0), warp, lane);
      }
    } else {
      if (n0 + 8 < l) block_reflect_s<BR>(S, Vtop + buf * 64, Tsm + buf * 64, c0, n0 + 8, l, warp - BQ_ROW_WARPS, BQ_WARPS - BQ_ROW_WARPS, lane);
    }
    __syncthreads();
  }
  double* Rb = Rstack + (size_t)blockIdx.x * l;
  for (int k = warp; k < l; k += BQ_WARPS) {
    double* dst = Y + (size_t)k * ldy + r0;
    for (int i = lane; i < BR; i += 32) {
      const double val = S[(size_t)k * LDS + i];
      if (i < nrows) dst[i] = val;
      if (i < l) Rb[(size_t)k * ldr + i] = (i <= k) ? val : 0.0;
    }
  }
  for (int j = threadIdx.x; j < l; j += BQ_THREADS) tau_g[(size_t)blockIdx.x * l + j] = tau_s[j];
}

// Up to 16 tree levels can be processed by one launch (their blocks are independent when every level starts from the
// identity): the block looks its level up in this table.
struct ApplyLevel { const double* V; long long ldv; long long rows; const double* Tg; const double* Ctop; long long ldc; double* Q; long long ldq; int first_block; };
struct ApplyTable { ApplyLevel lv[16]; int n; };

template <int BR>
__global__ void __launch_bounds__(BQ_THREADS, 1)
k_house_apply_blk(const ApplyTable tab, int l) {
  constexpr int LDS = BR + 4;
  constexpr int VPT = 8 * BR / BQ_THREADS;       // staged reflector values per thread
  extern __shared__ double sm[];
  double* S = sm;
  double* Vp = S + (size_t)l * LDS;
  double* Tsm = Vp + (size_t)8 * LDS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int li = 0;
  while (li + 1 < tab.n && (int)blockIdx.x >= tab.lv[li + 1].first_block) ++li;
  const double* V = tab.lv[li].V; const long long ldv = tab.lv[li].ldv; const long long rows = tab.lv[li].rows;
  const double* Tg = tab.lv[li].Tg; const double* Ctop = tab.lv[li].Ctop; const long long ldc = tab.lv[li].ldc;
  double* Q = tab.lv[li].Q; const long long ldq = tab.lv[li].ldq;
  const int blk = (int)blockIdx.x - tab.lv[li].first_block;
  const long long r0 = (long long)blk * BR;
  const int nrows = (int)min((long long)BR, rows - r0);
  const int npanels = (l + 7) / 8;

  for (int k = warp; k < l; k += BQ_WARPS) {
    for (int i = lane; i < BR; i += 32) {
      double c = 0.0;
      if (i < l) c = Ctop ? Ctop[(size_t)k * ldc + (size_t)blk * l + i] : (i == k ? 1.0 : 0.0);
      S[(size_t)k * LDS + i] = c;
    }
  }
  const double* Tblock = Tg + (size_t)blk * npanels * 64;
  // stage panel p: thread e handles elements e, e + 256, ... of the 8 x BR panel (column-major)
  double vreg[VPT];
  auto fetch = [&](int p) {
    const int c0 = 8 * p, pb = min(8, l - c0);
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int e = threadIdx.x + q * BQ_THREADS; const int c = e / BR, i = e % BR;
      const int d = c0 + c;
      double v = 0.0;
      if (c < pb) { if (i > d) v = (i < nrows) ? V[(size_t)d * ldv + r0 + i] : 0.0; else if (i == d) v = 1.0; }
      vreg[q] = v;
    }
  };
  fetch(npanels - 1);
  for (int p = npanels - 1; p >= 0; --p) {
    __syncthreads();                               // previous panel's readers are done with Vp / Tsm
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      const int e = threadIdx.x + q * BQ_THREADS; const int c = e / BR, i = e % BR;
      Vp[(size_t)c * LDS + i] = vreg[q];
    }
    if (threadIdx.x < 64) Tsm[threadIdx.x] = Tblock[(size_t)p * 64 + threadIdx.x];
    if (p > 0) fetch(p - 1);
    __syncthreads();
    block_reflect<BR>(S, Vp, Tsm, 8 * p, 0, l, warp, BQ_WARPS, lane, false);
  }
  __syncthreads();
  for (int k = warp; k < l; k += BQ_WARPS) {
    double* dst = Q + (size_t)k * ldq + r0;
    for (int i = lane; i < nrows; i += 32) dst[i] = S[(size_t)k * LDS + i];
  }
}

// Out (rows x l) = blockwise N * C:  rows [b*BR, (b+1)*BR) of N are multiplied by the l x l block C_b = rows [b*l, (b+1)*l) of
// Cpar (ldc).  One CTA per (node b, 64-row chunk).  Used to chain the explicit node factors of the upper tree levels.
constexpr int NG_ROWS = 64;
__global__ void __launch_bounds__(256)
k_node_gemm(const double* __restrict__ N, long long ldn, long long rows, int l, int BR, const double* __restrict__ Cpar, long long ldc,
            double* __restrict__ Out, long long ldo) {
  extern __shared__ double sm[];
  double* Cs = sm;                          // l x l, column-major, ld = l
  double* Ns = sm + (size_t)l * l;          // NG_ROWS x l, stored [k][row] with ld = NG_ROWS + 1
  const int chunks = BR / NG_ROWS;
  const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
  const long long r0 = (long long)b * BR + (long long)ch * NG_ROWS;
  if (r0 >= rows) return;
  const int nr = (int)min((long long)NG_ROWS, rows - r0);
  for (int e = threadIdx.x; e < l * l; e += 256) { const int i = e % l, k = e / l; Cs[e] = Cpar[(size_t)k * ldc + (size_t)b * l + i]; }
  for (int e = threadIdx.x; e < NG_ROWS * l; e += 256) {
    const int i = e % NG_ROWS, k = e / NG_ROWS;
    Ns[(size_t)k * (NG_ROWS + 1) + i] = (i < nr) ? N[(size_t)k * ldn + r0 + i] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // rows 4*ty .. 4*ty+3, columns tx + 16*j
  double acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int k = 0; k < l; ++k) {
    double a[4], bb[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = Ns[(size_t)k * (NG_ROWS + 1) + 4 * ty + i];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = tx + 16 * j; bb[j] = (c < l) ? Cs[(size_t)c * l + k] : 0.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = tx + 16 * j;
    if (c < l) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { const int r = 4 * ty + i; if (r < nr) Out[(size_t)c * ldo + r0 + r] = acc[i][j]; }
    }
  }
}

#include "tsqr_cluster.cuh"

// Fallback for panels wider than shared memory allows (l > 220 here): one CTA, unblocked Householder directly in global
// memory (L2-resident for the sizes that reach it).  Same outputs as k_house_factor with a single leaf.
__global__ void __launch_bounds__(1024, 1)
k_house_factor_global(double* Y, long long ldy, long long rows, int l, double* tau_g, double* R, long long ldr) {
  __shared__ double red[32];
  __shared__ double s_tau, s_beta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nref = (int)min((long long)l, rows);
  for (int j = 0; j < nref; ++j) {
    double* cj = Y + (size_t)j * ldy;
    double tail = 0.0;
    for (long long i = j + 1 + tid; i < rows; i += 1024) tail += cj[i] * cj[i];
    tail = warp_sum(tail);
    if (lane == 0) red[warp] = tail;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0; for (int w = 0; w < 32; ++w) t += red[w];
      const double x0 = cj[j];
      double beta, tau;
      if (t <= DBL_MIN) { tau = 0.0; beta = x0; }
      else { beta = sqrt(x0 * x0 + t); if (x0 >= 0.0) beta = -beta; tau = (beta - x0) / beta; }
      s_tau = tau; s_beta = beta; tau_g[j] = tau;
    }
    __syncthreads();
    const double tau = s_tau, beta = s_beta;
    const double x0 = cj[j];
    const double scale = (tau == 0.0) ? 0.0 : 1.0 / (x0 - beta);
    __syncthreads();
    for (long long i = j + 1 + tid; i < rows; i += 1024) cj[i] *= scale;
    if (tid == 0) cj[j] = beta;
    __syncthreads();
    if (tau != 0.0) {
      for (int k = j + 1 + warp; k < l; k += 32) {
        double* ck = Y + (size_t)k * ldy;
        double dot = 0.0;
        for (long long i = j + 1 + lane; i < rows; i += 32) dot += cj[i] * ck[i];
        dot = warp_sum(dot) + ck[j];
        const double w = tau * dot;
        for (long long i = j + 1 + lane; i < rows; i += 32) ck[i] -= w * cj[i];
        if (lane == 0) ck[j] -= w;
      }
    }
    __syncthreads();
  }
  for (int j = nref + tid; j < l; j += 1024) tau_g[j] = 0.0;
  for (long long e = tid; e < (long long)l * l; e += 1024) {
    const int i = (int)(e % l), k = (int)(e / l);
    R[(size_t)k * ldr + i] = (i <= k && i < nref) ? Y[(size_t)k * ldy + i] : 0.0;
  }
}
__global__ void __launch_bounds__(1024, 1)
k_house_apply_global(const double* V, long long ldv, long long rows, int nrefl, const double* tau_g, const double* Ctop, long long ldc,
                     double* Q, long long ldq, int l) {
  // Q must NOT alias V here.  Q (rows x l) = H_0 ... H_{nref-1} [C; 0];  C = Ctop (nrefl x l) or the identity.
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nref = (int)min((long long)nrefl, rows);
  for (long long e = tid; e < rows * l; e += 1024) {
    const long long i = e % rows; const int k = (int)(e / rows);
    double c = 0.0;
    if (Ctop) { if (i < nrefl) c = Ctop[(size_t)k * ldc + i]; } else c = (i == k ? 1.0 : 0.0);
    Q[(size_t)k * ldq + i] = c;
  }
  __syncthreads();
  for (int j = nref - 1; j >= 0; --j) {
    const double tau = tau_g[j];
    if (tau != 0.0) {
      const double* vj = V + (size_t)j * ldv;
      for (int k = warp; k < l; k += 32) {
        double* ck = Q + (size_t)k * ldq;
        double dot = 0.0;
        for (long long i = j + 1 + lane; i < rows; i += 32) dot += vj[i] * ck[i];
        dot = warp_sum(dot) + ck[j];
        const double w = tau * dot;
        for (long long i = j + 1 + lane; i < rows; i += 32) ck[i] -= w * vj[i];
        if (lane == 0) ck[j] -= w;
      }
    }
    __syncthreads();
  }
}

__global__ void k_copy_matrix(const double* __restrict__ src, long long lds, double* __restrict__ dst, long long ldd, long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) for (int k = blockIdx.y; k < cols; k += gridDim.y) dst[(size_t)k * ldd + i] = src[(size_t)k * lds + i];
}

int pick_br(int l) {
  // BR * l * 8 bytes (+ tau) must fit in ~220 KB of shared memory
  // and a leaf must be at least twice as tall as it is wide, or the tree does not shrink
  const long long budget = 220 * 1024 - 8LL * l;
  if ((long long)512 * l * 8 <= budget && 512 >= 2 * l) return 512;
  if ((long long)256 * l * 8 <= budget && 256 >= 2 * l) return 256;
  if ((long long)128 * l * 8 <= budget && 128 >= 2 * l) return 128;
  return 0;
}

int cl0_max_nodes() { static int m = -1; if (m < 0) { const char* e = getenv("RSVDB_TSQR_CL0_MAX"); m = e ? atoi(e) : 16; } return m; }
int cl_max_nodes() { static int m = -1; if (m < 0) { const char* e = getenv("RSVDB_TSQR_CL_MAX"); m = e ? atoi(e) : 99; } return m; }

size_t blk_smem_bytes(int l) { return ((size_t)(l + 8) * (256 + 4) + 64 + 720 + (size_t)l) * sizeof(double); }

cudaError_t set_attr_blk_once() {
  static DevOnce done;
  if (done.get()) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_house_apply_blk<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_house_factor_la<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_node_factor_cl<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_node_apply_cl<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_node_factor_cl<CL0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_node_apply_cl<CL0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  done.set();
  return cudaSuccess;
}

template <int BR> cudaError_t set_attr_once() {
  static DevOnce done;
  if (done.get()) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_house_factor<BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_house_apply<BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  done.set();
  return cudaSuccess;
}

}  // namespace

cudaError_t Tsqr::plan(long long rows, int l) {
  rows_ = rows; l_ = l; levels_.clear();
  br_ = pick_br(l);
  blk_ = false;
  if (l >= 1 && 2 * l <= 256 && blk_smem_bytes(l) <= 227 * 1024) { br_ = 256; blk_ = true; }
  size_t need = 0;
  long long r = rows;
  if (br_ == 0 || r <= 0) {           // single global-memory leaf
    Level L; L.rows = r; L.nb = 1; L.off_R = need; need += (size_t)l * l; L.off_tau = need; need += (size_t)l; L.off_T = need; L.off_E = need; L.cl = false; L.node_rows = 0;
    levels_.push_back(L);
    off_top_ = need; need += (size_t)l * l;
    off_scratch_ = need; need += (size_t)std::max<long long>(r, 1) * l;   // apply_global cannot run in place
  } else {
    const bool use_cl = blk_ && cl_factor_smem(l, CL0) <= 227 * 1024 && cl_apply_smem(l) <= 227 * 1024;
    for (;;) {
      Level L; L.rows = r;
      // leaves: 256-row single-CTA blocks (all SMs busy); upper levels and mid-sized panels: 1024-row cluster nodes
      // cluster nodes pay off when the level is latency-bound (few nodes); a level with hundreds of nodes is
      // throughput-bound and runs better as independent 256-row blocks
      const long long nb_cl = (r + CL_ROWS - 1) / CL_ROWS;
      const long long nb_cl0 = (r + CL0_ROWS - 1) / CL0_ROWS;
      L.cl = (use_cl && r > 256 && nb_cl <= cl_max_nodes() && (!levels_.empty() || r <= CL_ROWS)) ? CL : 0;
      // Level 0 as 2048-row nodes (8-CTA clusters) exactly where that removes a tree level: the R stack of the 8-CTA nodes fits ONE
      // 4-CTA node while the R stack of 256-row leaves would not (l = 100: 2621 < rows <= 20480, e.g. the 20000 x 100 panels of A^T Q:
      // 10 nodes -> 1 node instead of 79 leaves -> 8 nodes -> 1 node).  Measured (profiles/r02_tsqr_experiments.txt): an 8-CTA node
      // level costs ~317 us against 134 (leaves) + 219 (4-CTA nodes), so it only pays when a whole level disappears.
      const long long n_leaves = (r + br_ - 1) / br_;
      if (use_cl && levels_.empty() && r > CL_ROWS && nb_cl0 <= cl0_max_nodes() && n_leaves * l > CL_ROWS && nb_cl0 * l <= CL_ROWS) L.cl = CL0;
      L.node_rows = L.cl ? L.cl * 256 : br_;
      L.nb = (int)((r + L.node_rows - 1) / L.node_rows);
      L.off_R = need; need += (size_t)L.nb * l * l;
      L.off_tau = need; need += (size_t)L.nb * l;
      L.off_T = need; if (blk_) need += (size_t)L.nb * ((l + 7) / 8) * 64;
      L.off_E = need; if (blk_ && !levels_.empty()) need += (size_t)r * l;       // chained explicit factor of this (upper) level
      levels_.push_back(L);
      if (L.nb == 1) break;
      r = (long long)L.nb * l;
    }
    off_top_ = need; need += (size_t)l * l;
    off_scratch_ = need;
  }
  return ws_->reserve(need * sizeof(double));
}

cudaError_t Tsqr::factor(cudaStream_t st, double* Y, long long ldy, int* launches) {
  double* base = ws_->ptr;
  upper_done_ = false;
  if (br_ == 0) {
    Level& L = levels_[0];
    k_house_factor_global<<<1, 1024, 0, st>>>(Y, ldy, L.rows, l_, base + L.off_tau, base + L.off_R, l_);
    if (launches) ++*launches;
    return cudaGetLastError();
  }
  double* cur = Y; long long ld = ldy;
  for (size_t i = 0; i < levels_.size(); ++i) {
    Level& L = levels_[i];
    const long long ldr = (long long)L.nb * l_;
    const size_t smem = ((size_t)br_ * l_ + l_) * sizeof(double);
    cudaError_t e;
    if (blk_) {
      e = set_attr_blk_once(); if (e != cudaSuccess) return e;
      if (L.cl == CL0)
        k_node_factor_cl<CL0><<<L.nb * CL0, BQ_THREADS, cl_factor_smem(l_, CL0), st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr, base + L.off_T);
      else if (L.cl)
        k_node_factor_cl<CL><<<L.nb * CL, BQ_THREADS, cl_factor_smem(l_, CL), st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr, base + L.off_T);
      else
        k_house_factor_la<256><<<L.nb, BQ_THREADS, blk_smem_bytes(l_), st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr, base + L.off_T);
    } else
    switch (br_) {
      case 512: e = set_attr_once<512>(); if (e != cudaSuccess) return e;
        k_house_factor<512><<<L.nb, QR_THREADS, smem, st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr); break;
      case 256: e = set_attr_once<256>(); if (e != cudaSuccess) return e;
        k_house_factor<256><<<L.nb, QR_THREADS, smem, st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr); break;
      default: e = set_attr_once<128>(); if (e != cudaSuccess) return e;
        k_house_factor<128><<<L.nb, QR_THREADS, smem, st>>>(cur, ld, L.rows, l_, base + L.off_tau, base + L.off_R, ldr); break;
    }
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    if (blk_ && i >= 1 && side_ && ev_ && i < 16) {
      // level i is factored: form its explicit factor N_i = H [I; 0] (in place over the reflectors) on the side stream
      // while the next level is being factored on the main stream
      e = cudaEventRecord(ev_[i], st); if (e != cudaSuccess) return e;
      e = cudaStreamWaitEvent(side_, ev_[i], 0); if (e != cudaSuccess) return e;
      ApplyTable tab; tab.n = 1;
      tab.lv[0].V = cur; tab.lv[0].ldv = ld; tab.lv[0].rows = L.rows; tab.lv[0].Tg = base + L.off_T; tab.lv[0].Ctop = nullptr; tab.lv[0].ldc = 0;
      tab.lv[0].Q = cur; tab.lv[0].ldq = ld; tab.lv[0].first_block = 0;
      if (L.cl) k_node_apply_cl<CL><<<L.nb * CL, BQ_THREADS, cl_apply_smem(l_), side_>>>(tab, l_);      // upper levels are never CL0
      else k_house_apply_blk<256><<<L.nb, BQ_THREADS, blk_smem_bytes(l_), side_>>>(tab, l_);
      e = cudaGetLastError(); if (e != cudaSuccess) return e;
      if (launches) ++*launches;
      upper_done_ = true;
    }
    cur = base + L.off_R; ld = ldr;
  }
  if (upper_done_) { cudaError_t e = cudaEventRecord(ev_[16], side_); if (e != cudaSuccess) return e; }
  return cudaSuccess;
}

const double* Tsqr::R_local() const { return ws_->ptr + levels_.back().off_R; }
double* Tsqr::top_block() { return ws_->ptr + off_top_; }

cudaError_t Tsqr::form_q(cudaStream_t st, double* Y, long long ldy, const double* Ctop, long long ldc, int* launches) {
  double* base = ws_->ptr;
  if (br_ == 0) {
    Level& L = levels_[0];
    double* scratch = base + off_scratch_;
    k_house_apply_global<<<1, 1024, 0, st>>>(Y, ldy, L.rows, l_, base + L.off_tau, Ctop, ldc, scratch, L.rows, l_);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    if (L.rows > 0) {
      dim3 g((unsigned)((L.rows + 255) / 256), (unsigned)std::min(l_, 64));
      k_copy_matrix<<<g, 256, 0, st>>>(scratch, L.rows, Y, ldy, L.rows, l_);
      e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    if (launches) *launches += 2;
    return cudaSuccess;
  }
  if (blk_) {
    const int nlev = (int)levels_.size();
    const size_t smem = blk_smem_bytes(l_);
    auto level_V = [&](int i, double** V, long long* ldv) {
      if (i == 0) { *V = Y; *ldv = ldy; } else { *V = base + levels_[i - 1].off_R; *ldv = (long long)levels_[i - 1].nb * l_; }
    };
    // launch the levels listed in `which` (all of one kind) in one grid
    auto launch_apply = [&](const std::vector<int>& which, int cl, const double* C0, long long ldc0) -> cudaError_t {
      if (which.empty()) return cudaSuccess;
      if (which.size() > 16) return cudaErrorInvalidValue;
      ApplyTable tab; tab.n = 0; int first = 0;
      for (int i : which) {
        double* V; long long ldv; level_V(i, &V, &ldv);
        ApplyLevel& a = tab.lv[tab.n++];
        a.V = V; a.ldv = ldv; a.rows = levels_[i].rows; a.Tg = base + levels_[i].off_T; a.Ctop = C0; a.ldc = ldc0; a.Q = V; a.ldq = ldv;
        a.first_block = first; first += levels_[i].nb * (cl ? cl : 1);
      }
      if (cl == CL0) k_node_apply_cl<CL0><<<first, BQ_THREADS, cl_apply_smem(l_), st>>>(tab, l_);
      else if (cl) k_node_apply_cl<CL><<<first, BQ_THREADS, cl_apply_smem(l_), st>>>(tab, l_);
      else k_house_apply_blk<256><<<first, BQ_THREADS, smem, st>>>(tab, l_);
      if (launches) ++*launches;
      return cudaGetLastError();
    };
    const double* Cleaf = Ctop; long long ldcleaf = ldc;
    if (nlev > 1) {
      // (A) explicit factor N_i = H [I; 0] of every upper level (in place over the reflectors): one launch per node kind
      std::vector<int> up_cl, up_sc;
      for (int i = 1; i < nlev; ++i) (levels_[i].cl ? up_cl : up_sc).push_back(i);
      cudaError_t e = cudaSuccess;
      if (upper_done_) { e = cudaStreamWaitEvent(st, ev_[16], 0); if (e != cudaSuccess) return e; }   // formed on the side stream during factor()
      else {
        e = launch_apply(up_cl, CL, nullptr, 0); if (e != cudaSuccess) return e;
        e = launch_apply(up_sc, 0, nullptr, 0); if (e != cudaSuccess) return e;
      }
      // (B) chain top-down: E_i[node b] = N_i[node b] * E_{i+1}[rows b*l .. (b+1)*l)   (E_top = N_top * Ctop)
      static DevOnce ng_attr;
      const size_t ng_smem = ((size_t)l_ * l_ + (size_t)l_ * (NG_ROWS + 1)) * sizeof(double);
      if (!ng_attr.get()) { e = cudaFuncSetAttribute(k_node_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); if (e != cudaSuccess) return e; ng_attr.set(); }
      const double* Epar = Ctop; long long ldpar = ldc;
      for (int i = nlev - 1; i >= 1; --i) {
        double* V; long long ldv; level_V(i, &V, &ldv);
        if (Epar) {
          double* E = base + levels_[i].off_E;
          const int nr = levels_[i].node_rows;
          k_node_gemm<<<levels_[i].nb * (nr / NG_ROWS), 256, ng_smem, st>>>(V, ldv, levels_[i].rows, l_, nr, Epar, ldpar, E, levels_[i].rows);
          e = cudaGetLastError(); if (e != cudaSuccess) return e;
          if (launches) ++*launches;
          Epar = E; ldpar = levels_[i].rows;
        } else { Epar = V; ldpar = ldv; }                        // top level with the identity on top: E_top = N_top
      }
      Cleaf = Epar; ldcleaf = ldpar;
    }
    // (C) level 0
    return launch_apply(std::vector<int>{0}, levels_[0].cl, Cleaf, ldcleaf);
  }
  for (int i = (int)levels_.size() - 1; i >= 0; --i) {
    Level& L = levels_[i];
    double* V; long long ldv;
    if (i == 0) { V = Y; ldv = ldy; } else { V = base + levels_[i - 1].off_R; ldv = (long long)levels_[i - 1].nb * l_; }
    const double* C; long long ldcc;
    if (i == (int)levels_.size() - 1) { C = Ctop; ldcc = ldc; } else { C = base + L.off_R; ldcc = (long long)L.nb * l_; }
    const size_t smem = ((size_t)br_ * l_ + l_) * sizeof(double);
    switch (br_) {
      case 512: k_house_apply<512><<<L.nb, QR_THREADS, smem, st>>>(V, ldv, L.rows, l_, base + L.off_tau, C, ldcc, V, ldv); break;
      case 256: k_house_apply<256><<<L.nb, QR_THREADS, smem, st>>>(V, ldv, L.rows, l_, base + L.off_tau, C, ldcc, V, ldv); break;
      default:  k_house_apply<128><<<L.nb, QR_THREADS, smem, st>>>(V, ldv, L.rows, l_, base + L.off_tau, C, ldcc, V, ldv); break;
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    if (launches) ++*launches;
  }
  return cudaSuccess;
}


// Full QR for the reference's qr_decomposition_full / QRFullDecomposition API (O(rows^2) storage, small inputs only):
// F (rows x cols) is factored in place, R (rows x cols upper-trapezoidal) and Q (rows x rows) are written.
cudaError_t house_full_qr(cudaStream_t st, double* F, long long ldf, long long rows, int cols, double* tau, double* Rsq, double* Q,
                          long long ldq, int* launches) {
  k_house_factor_global<<<1, 1024, 0, st>>>(F, ldf, rows, cols, tau, Rsq, cols);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
  k_house_apply_global<<<1, 1024, 0, st>>>(F, ldf, rows, cols, tau, nullptr, 0, Q, ldq, (int)rows);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

}  // namespace rsvdb
