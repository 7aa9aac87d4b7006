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
#include "tsqr.cuh"

#include <algorithm>
#include <cfloat>

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

template <int BR> cudaError_t set_attr_once() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_house_factor<BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_house_apply<BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

}  // namespace

cudaError_t Tsqr::plan(long long rows, int l) {
  rows_ = rows; l_ = l; levels_.clear();
  br_ = pick_br(l);
  size_t need = 0;
  long long r = rows;
  if (br_ == 0 || r <= 0) {           // single global-memory leaf
    Level L; L.rows = r; L.nb = 1; L.off_R = need; need += (size_t)l * l; L.off_tau = need; need += (size_t)l;
    levels_.push_back(L);
    off_top_ = need; need += (size_t)l * l;
    off_scratch_ = need; need += (size_t)std::max<long long>(r, 1) * l;   // apply_global cannot run in place
  } else {
    for (;;) {
      Level L; L.rows = r; L.nb = (int)((r + br_ - 1) / br_);
      L.off_R = need; need += (size_t)L.nb * l * l;
      L.off_tau = need; need += (size_t)L.nb * l;
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
    cur = base + L.off_R; ld = ldr;
  }
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
