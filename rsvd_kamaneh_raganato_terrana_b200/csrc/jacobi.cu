// K6: SVD of the small k x k factor by one-sided (Hestenes) Jacobi with round-robin column pairing.
//
// Replaces the rotation loops of SVD<Jacobi>::jacobiSVD (reference include/SVD_class.hpp:125-156, rotations from
// src/JacobiOperations.cpp:6-103) and of SVD<ParallelJacobi>::ParallelJacobiSVD (:252-306), plus the abs / sort epilogue
// (:158-178).  The reference sweeps a two-sided (Kogbetliantz) Jacobi serially over (p,q); here the k x k work matrix
// (the R factor of the QR preconditioner, exactly as in the reference: :110-123) lives in one CTA's shared memory
// together with the accumulated right rotations, the k/2 disjoint column pairs of a round-robin step are rotated by
// different warps at once, the three inner products of a pair are reduced with warp shuffles, and a sweep is k-1 steps.
// Same singular values (to ~1e-15 relative, tighter than ParallelJacobi's own 1e-12 squared-weight stop), singular
// vectors equal up to sign / rotation inside clusters.  Output order and sign follow the reference: S descending,
// S >= 0 (:158-178).
#include "jacobi.cuh"

#include <cfloat>

namespace rsvdb {

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int JMAX_RPL = 32;   // rows per lane of a 16-lane group: k <= 512

// X, Z: k x k column-major with leading dimension k, in shared or global memory.
// One column pair per HALF warp (16 lanes, RPL rows per lane): with 1024 threads all k/2 <= 64 pairs of a round-robin
// step rotate at once.  The rotation angle comes from two rsqrt (no FP64 division or square root on the critical path):
//   d = b - a, r = sqrt(d^2 + 4 g^2), cos(2t) = |d| / r, c = sqrt((1 + cos 2t) / 2), s = sign(d) g / (r c).
template <int RPL>
__global__ void __launch_bounds__(1024, 1)
k_jacobi(const double* __restrict__ W, long long ldw, int k, int transpose_in, double* __restrict__ Uo, long long ldu,
         double* __restrict__ So, double* __restrict__ Zo, long long ldz, double* Xg, double* Zg, int use_smem,
         int max_sweeps, int* __restrict__ info) {
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ int s_total;
  double* X = use_smem ? sm : Xg;
  double* Z = use_smem ? sm + (size_t)k * k : Zg;
  double* sig = use_smem ? sm + 2 * (size_t)k * k : Zg + (size_t)k * k;   // k doubles
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int hl = tid & 15, grp = tid >> 4, ngrp = blockDim.x >> 4;         // half-warp lane / group

  for (int e = tid; e < k * k; e += blockDim.x) {
    const int i = e % k, j = e / k;
    X[e] = transpose_in ? W[(size_t)i * ldw + j] : W[(size_t)j * ldw + i];
    Z[e] = (i == j) ? 1.0 : 0.0;
  }
  if (tid == 0) s_total = 0;
  __syncthreads();

  const int n = (k + 1) & ~1;          // players in the round-robin (one dummy when k is odd)
  const int npairs = n >> 1;
  const double tol = sqrt((double)k) * (0.5 * DBL_EPSILON);
  const double tol2 = tol * tol;
  int sweep = 0; bool converged = (k < 2);
  while (!converged && sweep < max_sweeps) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int step = 0; step < n - 1; ++step) {
      int nrot = 0;
      for (int pi = grp; pi < ((npairs + ngrp - 1) / ngrp) * ngrp; pi += ngrp) {   // uniform trip count per half warp pair
        int p = 0, q = k;
        if (pi < npairs) {
          if (pi == 0) { p = n - 1; q = step; }
          else { p = (step + pi) % (n - 1); q = (step - pi + (n - 1)) % (n - 1); }
          if (p > q) { const int t = p; p = q; q = t; }
        }
        const bool live = q < k;        // not the dummy player, not a padding pair
        double* xp = X + (size_t)(live ? p : 0) * k; double* xq = X + (size_t)(live ? q : 0) * k;
        double a = 0.0, b = 0.0, g = 0.0, vp[RPL], vq[RPL];
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
          const int i = hl + 16 * ii;
          const bool on = live && i < k;
          vp[ii] = on ? xp[i] : 0.0; vq[ii] = on ? xq[i] : 0.0;
          a = fma(vp[ii], vp[ii], a); b = fma(vq[ii], vq[ii], b); g = fma(vp[ii], vq[ii], g);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (live && g * g > tol2 * (a * b) && fabs(g) > DBL_MIN) {
          ++nrot;
          const double d = b - a;
          const double n2 = fma(d, d, 4.0 * g * g);
          const double ir = rsqrt(n2);                       // 1 / r
          const double c2 = fma(0.5 * fabs(d), ir, 0.5);     // cos^2 = (1 + |d| / r) / 2   in [0.5, 1]
          const double ic = rsqrt(c2);
          const double c = c2 * ic;
          const double s = ((d >= 0.0) ? g : -g) * ir * ic;
          double* zp = Z + (size_t)p * k; double* zq = Z + (size_t)q * k;
#pragma unroll
          for (int ii = 0; ii < RPL; ++ii) {
            const int i = hl + 16 * ii;
            if (i < k) {
              xp[i] = c * vp[ii] - s * vq[ii];
              xq[i] = s * vp[ii] + c * vq[ii];
              const double z1 = zp[i], z2 = zq[i];
              zp[i] = c * z1 - s * z2;
              zq[i] = s * z1 + c * z2;
            }
          }
        }
      }
      if (nrot && hl == 0) atomicAdd(&s_rot, nrot);
      __syncthreads();
    }
    ++sweep;
    converged = (s_rot == 0);
    if (tid == 0) s_total += s_rot;
    __syncthreads();
  }

  // singular values = column norms
  for (int j = warp; j < k; j += nwarps) {
    double a = 0.0;
    for (int i = lane; i < k; i += 32) a = fma(X[(size_t)j * k + i], X[(size_t)j * k + i], a);
    a = wsum(a);
    if (lane == 0) sig[j] = sqrt(a);
  }
  __syncthreads();
  // rank sort, descending; ties broken by index so the permutation is deterministic
  for (int j = warp; j < k; j += nwarps) {
    const double sj = sig[j];
    int r = 0;
    for (int i = lane; i < k; i += 32) { const double si = sig[i]; r += (si > sj || (si == sj && i < j)) ? 1 : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
    if (lane == 0) So[r] = sj;
    for (int i = lane; i < k; i += 32) {
      Uo[(size_t)r * ldu + i] = (sj > 0.0) ? X[(size_t)j * k + i] * inv : (i == j ? 1.0 : 0.0);
      Zo[(size_t)r * ldz + i] = Z[(size_t)j * k + i];
    }
  }
  if (tid == 0 && info) { info[0] = converged ? sweep : -sweep; info[1] = s_total; }
}

}  // namespace

cudaError_t jacobi_svd_square(GemmWorkspace& ws, cudaStream_t st, const double* W, long long ldw, int k, int transpose_in,
                              double* U, long long ldu, double* S, double* Z, long long ldz, int* d_info, int* launches) {
  if (k <= 0) return cudaSuccess;
  if (k > 16 * JMAX_RPL) return cudaErrorInvalidValue;
  const size_t smem_need = (2 * (size_t)k * k + k) * sizeof(double);
  const int use_smem = smem_need <= 220 * 1024;
  double* Xg = nullptr; double* Zg = nullptr;
  if (!use_smem) {
    cudaError_t e = ws.reserve(smem_need + 64); if (e != cudaSuccess) return e;
    Xg = ws.ptr; Zg = ws.ptr + (size_t)k * k;
  }
  const size_t smem = use_smem ? smem_need : 0;
  const int threads = k >= 48 ? 1024 : (k >= 24 ? 512 : 256);
  const int rpl = (k + 15) / 16;
#define JLAUNCH(R)                                                                                                   \
  { static bool attr = false;                                                                                        \
    if (!attr) { cudaError_t e = cudaFuncSetAttribute(k_jacobi<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); \
                 if (e != cudaSuccess) return e; attr = true; }                                                      \
    k_jacobi<R><<<1, threads, smem, st>>>(W, ldw, k, transpose_in, U, ldu, S, Z, ldz, Xg, Zg, use_smem, 60, d_info); }
  if (rpl <= 1) JLAUNCH(1) else if (rpl <= 2) JLAUNCH(2) else if (rpl <= 4) JLAUNCH(4) else if (rpl <= 7) JLAUNCH(7) else if (rpl <= 8) JLAUNCH(8)
  else if (rpl <= 16) JLAUNCH(16) else JLAUNCH(32)
#undef JLAUNCH
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace rsvdb
