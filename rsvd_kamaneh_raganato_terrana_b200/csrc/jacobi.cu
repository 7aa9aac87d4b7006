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
#include "dev_once.cuh"
#include "jacobi.cuh"
#include "ptx.cuh"

#include <cfloat>
#include <cstdlib>

namespace rsvdb {

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int JMAX_RPL = 32;   // rows per lane of a 16-lane group: k <= 512
int grid_min() { static int m = -1; if (m < 0) { const char* e = getenv("RSVDB_JACOBI_GRID_MIN"); m = e ? atoi(e) : 512; } return m; }

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
  // one pair per half warp and step (the launch guarantees npairs <= blockDim.x / 16); the round-robin seats are advanced
  // incrementally -- the two integer modulos per pair and step they replace were a quarter of the instruction stream
  const int pi = grp;
  while (!converged && sweep < max_sweeps) {
    if (tid == 0) s_rot = 0;
    int pr = (pi == 0) ? n - 1 : pi % (n - 1);
    int qr = (pi == 0) ? 0 : (n - 1 - pi % (n - 1)) % (n - 1);
    __syncthreads();
    for (int step = 0; step < n - 1; ++step) {
      int nrot = 0;
      {
        int p = 0, q = k;
        if (pi < npairs) { p = min(pr, qr); q = max(pr, qr); }
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
        if (pi != 0) pr = (pr + 1 == n - 1) ? 0 : pr + 1;
        qr = (qr + 1 == n - 1) ? 0 : qr + 1;
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

// ------------------------------------------------------------------------------------------------------------------
// Cluster variant: the single-CTA kernel above is bound by shared-memory bandwidth (every round-robin step reads and
// writes all of X and Z through one SM: 4 * k^2 * 8 bytes).  Here the ROWS of X and Z are split over a cluster of
// JC = 4 CTAs, so each SM moves a quarter of the bytes; per step and pair only the three partial inner products cross the
// cluster (remote st.shared::cluster of 3 doubles to each peer + one hardware cluster barrier).  Every CTA sums the
// partials in rank order and therefore takes bit-identical rotation decisions, which keeps the loops (and barriers) of
// the four CTAs in lock step without any further communication.
// ------------------------------------------------------------------------------------------------------------------
constexpr int JC = 4;

__device__ __forceinline__ unsigned jc_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void jc_sync() {
  // the .aligned forms need the whole warp converged; lanes of one warp take different branches in the pair loops
  // (a dummy pair for odd k, idle 16-lane groups), so reconverge explicitly first
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void jc_store(double* p, unsigned rank, double v) {
  unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(p)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

__device__ __forceinline__ unsigned jc_map(const void* p, unsigned rank) {
  unsigned la = static_cast<unsigned>(__cvta_generic_to_shared(p)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
// remote 8-byte store that credits 8 bytes to CTA `rank`'s copy of *bar (SASS STAS.64)
__device__ __forceinline__ void jc_st_async(double* p, unsigned rank, double v, uint64_t* bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(jc_map(p, rank)), "l"(__double_as_longlong(v)), "r"(jc_map(bar, rank)) : "memory");
}
__device__ __forceinline__ void jc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// RPL: rows of the CTA's slice per lane of a 16-lane group (slice rows sr = ceil(k / JC) <= 16 * RPL).
template <int RPL>
__global__ void __cluster_dims__(JC, 1, 1) __launch_bounds__(1024, 1)
k_jacobi_cl(const double* __restrict__ W, long long ldw, int k, int transpose_in, double* __restrict__ Uo, long long ldu,
            double* __restrict__ So, double* __restrict__ Zo, long long ldz, int max_sweeps, int* __restrict__ info) {
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ int s_total;
  const unsigned rank = jc_rank();
  const int sr = (k + JC - 1) / JC;                 // slice rows
  const int r0 = (int)rank * sr;
  const int nr = max(0, min(sr, k - r0));
  const int n = (k + 1) & ~1, npairs = n >> 1;
  double* X = sm;                                   // k columns of sr rows
  double* Z = X + (size_t)k * sr;
  double* xch = Z + (size_t)k * sr;                 // [2][npairs][JC][4]
  double* sig = xch + 2 * (size_t)npairs * JC * 4;  // [k][JC] partial squared norms, then [k] totals behind it
  const int tid = threadIdx.x, hl = tid & 15, grp = tid >> 4, ngrp = blockDim.x >> 4;

  for (int e = tid; e < k * sr; e += blockDim.x) {
    const int i = e % sr, j = e / sr, gi = r0 + i;
    double x = 0.0;
    if (i < nr) x = transpose_in ? W[(size_t)gi * ldw + j] : W[(size_t)j * ldw + gi];
    X[e] = x;
    Z[e] = (i < nr && gi == j) ? 1.0 : 0.0;
  }
  if (tid == 0) s_total = 0;
  __syncthreads();

  const double tol = sqrt((double)k) * (0.5 * DBL_EPSILON);
  const double tol2 = tol * tol;
  const int rounds = (npairs + ngrp - 1) / ngrp;
  int sweep = 0, buf = 0; bool converged = (k < 2);
  while (!converged && sweep < max_sweeps) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int step = 0; step < n - 1; ++step) {
      // (1) partial inner products of every pair over this CTA's rows -> all CTAs
      for (int rd = 0; rd < rounds; ++rd) {
        const int pi = grp + rd * ngrp;
        int p = 0, q = k;
        if (pi < npairs) {
          if (pi == 0) { p = n - 1; q = step; }
          else { p = (step + pi) % (n - 1); q = (step - pi + (n - 1)) % (n - 1); }
          if (p > q) { const int t = p; p = q; q = t; }
        }
        const bool live = q < k;
        const double* xp = X + (size_t)(live ? p : 0) * sr; const double* xq = X + (size_t)(live ? q : 0) * sr;
        double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
        for (int ii = 0; ii < RPL; ++ii) {
          const int i = hl + 16 * ii;
          const bool on = live && i < nr;
          const double vp = on ? xp[i] : 0.0, vq = on ? xq[i] : 0.0;
          a = fma(vp, vp, a); b = fma(vq, vq, b); g = fma(vp, vq, g);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (pi < npairs && hl < 3 * JC) {             // lanes 0..11: value hl % 3 to CTA hl / 3
          const double v = (hl % 3 == 0) ? a : ((hl % 3 == 1) ? b : g);
          jc_store(xch + (((size_t)buf * npairs + pi) * JC + rank) * 4 + (hl % 3), (unsigned)(hl / 3), v);
        }
      }
      jc_sync();
      // (2) every CTA takes the same decision and rotates its rows
      int nrot = 0;
      for (int rd = 0; rd < rounds; ++rd) {
        const int pi = grp + rd * ngrp;
        if (pi >= npairs) continue;
        int p, q;
        if (pi == 0) { p = n - 1; q = step; }
        else { p = (step + pi) % (n - 1); q = (step - pi + (n - 1)) % (n - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        if (q >= k) continue;
        const double* xx = xch + ((size_t)buf * npairs + pi) * JC * 4;
        double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
        for (int sR = 0; sR < JC; ++sR) { a += xx[sR * 4 + 0]; b += xx[sR * 4 + 1]; g += xx[sR * 4 + 2]; }
        if (g * g > tol2 * (a * b) && fabs(g) > DBL_MIN) {
          ++nrot;
          const double d = b - a;
          const double n2 = fma(d, d, 4.0 * g * g);
          const double ir = rsqrt(n2);
          const double c2 = fma(0.5 * fabs(d), ir, 0.5);
          const double ic = rsqrt(c2);
          const double c = c2 * ic;
          const double s = ((d >= 0.0) ? g : -g) * ir * ic;
          double* xp = X + (size_t)p * sr; double* xq = X + (size_t)q * sr;
          double* zp = Z + (size_t)p * sr; double* zq = Z + (size_t)q * sr;
#pragma unroll
          for (int ii = 0; ii < RPL; ++ii) {
            const int i = hl + 16 * ii;
            if (i < nr) {
              const double x1 = xp[i], x2 = xq[i];
              xp[i] = c * x1 - s * x2; xq[i] = s * x1 + c * x2;
              const double z1 = zp[i], z2 = zq[i];
              zp[i] = c * z1 - s * z2; zq[i] = s * z1 + c * z2;
            }
          }
        }
      }
      if (nrot && hl == 0) atomicAdd(&s_rot, nrot);
      __syncthreads();
      buf ^= 1;
    }
    ++sweep;
    converged = (s_rot == 0);                       // identical on every CTA: same sums, same decisions
    if (tid == 0) s_total += s_rot;
    __syncthreads();
  }

  // singular values: column norms summed over the cluster
  double* part = sig;                               // [k][JC]
  double* tot = sig + (size_t)k * JC;               // [k]
  // (the two 16-lane groups of a warp own different columns: keep the trip count uniform per warp, the shuffles are full-warp)
  for (int j0 = 0; j0 < k; j0 += ngrp) {
    const int j = j0 + grp; const bool on = j < k;
    double a = 0.0;
    if (on) for (int i = hl; i < nr; i += 16) a = fma(X[(size_t)j * sr + i], X[(size_t)j * sr + i], a);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (on && hl < JC) jc_store(part + (size_t)j * JC + rank, (unsigned)hl, a);
  }
  jc_sync();
  for (int j = tid; j < k; j += blockDim.x) {
    double a = 0.0;
#pragma unroll
    for (int sR = 0; sR < JC; ++sR) a += part[(size_t)j * JC + sR];
    tot[j] = sqrt(a);
  }
  __syncthreads();
  for (int j0 = 0; j0 < k; j0 += ngrp) {
    const int j = j0 + grp; const bool on = j < k;
    const double sj = on ? tot[j] : 0.0;
    int r = 0;
    if (on) for (int i = hl; i < k; i += 16) { const double si = tot[i]; r += (si > sj || (si == sj && i < j)) ? 1 : 0; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (!on) continue;
    const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
    if (rank == 0 && hl == 0) So[r] = sj;
    for (int i = hl; i < nr; i += 16) {
      const int gi = r0 + i;
      Uo[(size_t)r * ldu + gi] = (sj > 0.0) ? X[(size_t)j * sr + i] * inv : (gi == j ? 1.0 : 0.0);
      Zo[(size_t)r * ldz + gi] = Z[(size_t)j * sr + i];
    }
  }
  if (rank == 0 && tid == 0 && info) { info[0] = converged ? sweep : -sweep; info[1] = s_total; }
  jc_sync();                                        // peers may still be reading this CTA's shared memory
}

// Round-2 revision of the cluster kernel's step.  ncu on the version above: issue slots 50 % busy, barrier 3.8 + membar 1.9 stall
// cycles per issue -- every step paid one hardware cluster barrier (4096 threads, release/acquire fences over all memory) plus ~8
// integer modulos per thread for the round-robin pairing.  Here
//  * a pair's three partial inner products travel with st.async ... mbarrier::complete_tx to one mbarrier PER PAIR AND BUFFER on
//    every CTA, and only the half warp that owns the pair waits for it (no cluster barrier in the sweep at all; the block barrier
//    at the end of a step, which hands the rotated columns to their next owners, stays);
//  * the pairing is advanced incrementally (p, q -> p + 1, q + 1 mod n - 1) and kept in registers across both halves of a step.
// Buffer reuse is safe with two buffers: a CTA can pass the wait of step s + 1 only after every peer has sent its step-(s + 1)
// partials, i.e. after that peer's half warp has finished reading step s, so writes of step s + 2 never meet reads of step s.
template <int RPL, int ROUNDS>
__global__ void __cluster_dims__(JC, 1, 1) __launch_bounds__(1024, 1)
k_jacobi_cl2(const double* __restrict__ W, long long ldw, int k, int transpose_in, double* __restrict__ Uo, long long ldu,
             double* __restrict__ So, double* __restrict__ Zo, long long ldz, int max_sweeps, int* __restrict__ info) {
  extern __shared__ double sm[];
  __shared__ int s_rot;
  __shared__ int s_total;
  const unsigned rank = jc_rank();
  const int sr = (k + JC - 1) / JC;                 // slice rows
  const int r0 = (int)rank * sr;
  const int nr = max(0, min(sr, k - r0));
  const int n = (k + 1) & ~1, npairs = n >> 1;
  double* X = sm;                                   // k columns of sr rows
  double* Z = X + (size_t)k * sr;
  double* xch = Z + (size_t)k * sr;                 // [2][npairs][JC][4]
  double* sig = xch + 2 * (size_t)npairs * JC * 4;  // [k][JC] partial squared norms, then [k] totals behind it
  uint64_t* bar = reinterpret_cast<uint64_t*>(sig + (size_t)k * JC + k);   // [2][npairs]
  const int tid = threadIdx.x, hl = tid & 15, grp = tid >> 4, ngrp = blockDim.x >> 4;
  const unsigned hmask = (tid & 16) ? 0xffff0000u : 0x0000ffffu;   // the two half warps of a warp may own a live and a padding pair

  for (int e = tid; e < k * sr; e += blockDim.x) {
    const int i = e % sr, j = e / sr, gi = r0 + i;
    double x = 0.0;
    if (i < nr) x = transpose_in ? W[(size_t)gi * ldw + j] : W[(size_t)j * ldw + gi];
    X[e] = x;
    Z[e] = (i < nr && gi == j) ? 1.0 : 0.0;
  }
  for (int e = tid; e < 2 * npairs; e += blockDim.x) mbar_init(&bar[e], 1);
  if (tid == 0) s_total = 0;
  fence_mbar_init();
  __syncthreads();
  constexpr uint32_t TX = JC * 3 * sizeof(double);  // three doubles from every CTA of the cluster (this one included)
#pragma unroll
  for (int rd = 0; rd < ROUNDS; ++rd) {
    const int pi = grp + rd * ngrp;
    if (pi < npairs && hl == 0) { mbar_expect_tx(&bar[pi], TX); mbar_expect_tx(&bar[npairs + pi], TX); }
  }
  jc_sync();                                        // every CTA's barriers exist and are armed before the first remote store

  const double tol = sqrt((double)k) * (0.5 * DBL_EPSILON);
  const double tol2 = tol * tol;
  uint32_t phase = 0;                               // bit b: parity to wait for on buffer b
  int sweep = 0, buf = 0; bool converged = (k < 2);
  while (!converged && sweep < max_sweeps) {
    if (tid == 0) s_rot = 0;
    int pr[ROUNDS], qr[ROUNDS];                     // raw round-robin seats of this half warp's pairs at step 0
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int pi = grp + rd * ngrp;
      pr[rd] = (pi == 0) ? n - 1 : pi % (n - 1);
      qr[rd] = (pi == 0) ? 0 : (n - 1 - pi % (n - 1)) % (n - 1);
    }
    __syncthreads();
    for (int step = 0; step < n - 1; ++step) {
      int pp[ROUNDS], qq[ROUNDS];
      // (1) partial inner products of every pair over this CTA's rows -> all CTAs
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int pi = grp + rd * ngrp;
        const int p = min(pr[rd], qr[rd]), q = max(pr[rd], qr[rd]);
        pp[rd] = p; qq[rd] = q;
        if (pi < npairs) {
          const bool live = q < k;
          const double* xp = X + (size_t)(live ? p : 0) * sr; const double* xq = X + (size_t)(live ? q : 0) * sr;
          double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
          for (int ii = 0; ii < RPL; ++ii) {
            const int i = hl + 16 * ii;
            const bool on = live && i < nr;
            const double vp = on ? xp[i] : 0.0, vq = on ? xq[i] : 0.0;
            a = fma(vp, vp, a); b = fma(vq, vq, b); g = fma(vp, vq, g);
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            a += __shfl_xor_sync(hmask, a, o); b += __shfl_xor_sync(hmask, b, o); g += __shfl_xor_sync(hmask, g, o);
          }
          if (hl < 3 * JC) {                          // lanes 0..11: value hl % 3 to CTA hl / 3
            const double v = (hl % 3 == 0) ? a : ((hl % 3 == 1) ? b : g);
            jc_st_async(xch + (((size_t)buf * npairs + pi) * JC + rank) * 4 + (hl % 3), (unsigned)(hl / 3), v, &bar[buf * npairs + pi]);
          }
        }
      }
      // (2) every CTA takes the same decision and rotates its rows
      int nrot = 0;
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int pi = grp + rd * ngrp;
        const int p = pp[rd], q = qq[rd];
        if (pi < npairs) {
          jc_mbar_wait(&bar[buf * npairs + pi], (phase >> buf) & 1u);
          const double* xx = xch + ((size_t)buf * npairs + pi) * JC * 4;
          double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
          for (int sR = 0; sR < JC; ++sR) { a += xx[sR * 4 + 0]; b += xx[sR * 4 + 1]; g += xx[sR * 4 + 2]; }
          __syncwarp(hmask);     // all 16 lanes hold the sums before the barrier is re-armed
          if (hl == 0) mbar_expect_tx(&bar[buf * npairs + pi], TX);     // its next use is two steps away
          if (q < k && g * g > tol2 * (a * b) && fabs(g) > DBL_MIN) {
            ++nrot;
            const double d = b - a;
            const double n2 = fma(d, d, 4.0 * g * g);
            const double ir = rsqrt(n2);
            const double c2 = fma(0.5 * fabs(d), ir, 0.5);
            const double ic = rsqrt(c2);
            const double c = c2 * ic;
            const double s = ((d >= 0.0) ? g : -g) * ir * ic;
            double* xp = X + (size_t)p * sr; double* xq = X + (size_t)q * sr;
            double* zp = Z + (size_t)p * sr; double* zq = Z + (size_t)q * sr;
#pragma unroll
            for (int ii = 0; ii < RPL; ++ii) {
              const int i = hl + 16 * ii;
              if (i < nr) {
                const double x1 = xp[i], x2 = xq[i];
                xp[i] = c * x1 - s * x2; xq[i] = s * x1 + c * x2;
                const double z1 = zp[i], z2 = zq[i];
                zp[i] = c * z1 - s * z2; zq[i] = s * z1 + c * z2;
              }
            }
          }
        }
        // next step's seats: everybody but seat n - 1 moves up by one
        if (pi != 0) { pr[rd] = (pr[rd] + 1 == n - 1) ? 0 : pr[rd] + 1; }
        qr[rd] = (qr[rd] + 1 == n - 1) ? 0 : qr[rd] + 1;
      }
      if (nrot && hl == 0) atomicAdd(&s_rot, nrot);
      phase ^= (1u << buf);
      __syncthreads();
      buf ^= 1;
    }
    ++sweep;
    converged = (s_rot == 0);                       // identical on every CTA: same sums, same decisions
    if (tid == 0) s_total += s_rot;
    __syncthreads();
  }

  // singular values: column norms summed over the cluster
  double* part = sig;                               // [k][JC]
  double* tot = sig + (size_t)k * JC;               // [k]
  jc_sync();                                        // nobody is still inside the sweeps when the remote stores below land
  for (int j0 = 0; j0 < k; j0 += ngrp) {
    const int j = j0 + grp; const bool on = j < k;
    double a = 0.0;
    if (on) for (int i = hl; i < nr; i += 16) a = fma(X[(size_t)j * sr + i], X[(size_t)j * sr + i], a);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (on && hl < JC) jc_store(part + (size_t)j * JC + rank, (unsigned)hl, a);
  }
  jc_sync();
  for (int j = tid; j < k; j += blockDim.x) {
    double a = 0.0;
#pragma unroll
    for (int sR = 0; sR < JC; ++sR) a += part[(size_t)j * JC + sR];
    tot[j] = sqrt(a);
  }
  __syncthreads();
  for (int j0 = 0; j0 < k; j0 += ngrp) {
    const int j = j0 + grp; const bool on = j < k;
    const double sj = on ? tot[j] : 0.0;
    int r = 0;
    if (on) for (int i = hl; i < k; i += 16) { const double si = tot[i]; r += (si > sj || (si == sj && i < j)) ? 1 : 0; }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (!on) continue;
    const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
    if (rank == 0 && hl == 0) So[r] = sj;
    for (int i = hl; i < nr; i += 16) {
      const int gi = r0 + i;
      Uo[(size_t)r * ldu + gi] = (sj > 0.0) ? X[(size_t)j * sr + i] * inv : (gi == j ? 1.0 : 0.0);
      Zo[(size_t)r * ldz + gi] = Z[(size_t)j * sr + i];
    }
  }
  if (rank == 0 && tid == 0 && info) { info[0] = converged ? sweep : -sweep; info[1] = s_total; }
  jc_sync();                                        // peers may still be reading this CTA's shared memory
}

// ------------------------------------------------------------------------------------------------------------------
// Grid variant for factors that do not fit one SM (k > 512; the reference's SVD<Jacobi> / PCA<method> take any size,
// include/SVD_class.hpp:101-180).  X and Z (k x k each) live in global memory and stay L2-resident (16 MB at k = 1000);
// the k/2 disjoint column pairs of a round-robin step are independent, so one CTA rotates one pair and a step is one
// launch.  Same rotation formula, threshold and epilogue as k_jacobi; the sweep count is read back once per sweep.
// ------------------------------------------------------------------------------------------------------------------
constexpr int JG_THREADS = 128;

__global__ void k_jg_init(const double* __restrict__ W, long long ldw, int k, int transpose_in, double* __restrict__ X, double* __restrict__ Z) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)k * k) return;
  const int i = (int)(e % k), j = (int)(e / k);
  X[e] = transpose_in ? W[(size_t)i * ldw + j] : W[(size_t)j * ldw + i];
  Z[e] = (i == j) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(JG_THREADS)
k_jg_step(double* __restrict__ X, double* __restrict__ Z, int k, int step, double tol2, int* __restrict__ rot) {
  __shared__ double red[3][JG_THREADS / 32];
  const int n = (k + 1) & ~1;
  const int pi = blockIdx.x;
  int p, q;
  if (pi == 0) { p = n - 1; q = step; }
  else { p = (step + pi) % (n - 1); q = (step - pi + (n - 1)) % (n - 1); }
  if (p > q) { const int t = p; p = q; q = t; }
  if (q >= k) return;                                   // the dummy player of an odd k
  double* xp = X + (size_t)p * k; double* xq = X + (size_t)q * k;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int i = threadIdx.x; i < k; i += JG_THREADS) {
    const double vp = xp[i], vq = xq[i];
    a = fma(vp, vp, a); b = fma(vq, vq, b); g = fma(vp, vq, g);
  }
  a = wsum(a); b = wsum(b); g = wsum(g);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = g; }
  __syncthreads();
  a = 0.0; b = 0.0; g = 0.0;
#pragma unroll
  for (int w = 0; w < JG_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; g += red[2][w]; }   // same order in every thread
  if (!(g * g > tol2 * (a * b) && fabs(g) > DBL_MIN)) return;
  if (threadIdx.x == 0) atomicAdd(rot, 1);
  const double d = b - a;
  const double n2 = fma(d, d, 4.0 * g * g);
  const double ir = rsqrt(n2);
  const double c2 = fma(0.5 * fabs(d), ir, 0.5);
  const double ic = rsqrt(c2);
  const double c = c2 * ic;
  const double s = ((d >= 0.0) ? g : -g) * ir * ic;
  double* zp = Z + (size_t)p * k; double* zq = Z + (size_t)q * k;
  for (int i = threadIdx.x; i < k; i += JG_THREADS) {
    const double vp = xp[i], vq = xq[i];
    xp[i] = c * vp - s * vq; xq[i] = s * vp + c * vq;
    const double z1 = zp[i], z2 = zq[i];
    zp[i] = c * z1 - s * z2; zq[i] = s * z1 + c * z2;
  }
}

__global__ void k_jg_norms(const double* __restrict__ X, int k, double* __restrict__ sig) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= k) return;
  double a = 0.0;
  for (int i = lane; i < k; i += 32) a = fma(X[(size_t)j * k + i], X[(size_t)j * k + i], a);
  a = wsum(a);
  if (lane == 0) sig[j] = sqrt(a);
}

// rank sort (descending, ties by index) and the permuted write of U = X / sigma, Z  -- one CTA per column
__global__ void __launch_bounds__(256)
k_jg_finish(const double* __restrict__ X, const double* __restrict__ Z, const double* __restrict__ sig, int k, double* __restrict__ Uo, long long ldu,
            double* __restrict__ So, double* __restrict__ Zo, long long ldz, int* __restrict__ info, int sweeps_signed, int rotations) {
  __shared__ int cnt[8];
  const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double sj = sig[j];
  int r = 0;
  for (int i = threadIdx.x; i < k; i += 256) { const double si = sig[i]; r += (si > sj || (si == sj && i < j)) ? 1 : 0; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  if (lane == 0) cnt[warp] = r;
  __syncthreads();
  r = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) r += cnt[w];
  const double inv = (sj > 0.0) ? 1.0 / sj : 0.0;
  if (threadIdx.x == 0) So[r] = sj;
  for (int i = threadIdx.x; i < k; i += 256) {
    Uo[(size_t)r * ldu + i] = (sj > 0.0) ? X[(size_t)j * k + i] * inv : (i == j ? 1.0 : 0.0);
    Zo[(size_t)r * ldz + i] = Z[(size_t)j * k + i];
  }
  if (j == 0 && threadIdx.x == 0 && info) { info[0] = sweeps_signed; info[1] = rotations; }
}

cudaError_t jacobi_svd_grid(GemmWorkspace& ws, cudaStream_t st, const double* W, long long ldw, int k, int transpose_in,
                            double* U, long long ldu, double* S, double* Z, long long ldz, int* d_info, int max_sweeps, int* launches) {
  const size_t kk = (size_t)k * k;
  cudaError_t e = ws.reserve((2 * kk + k + 16) * sizeof(double)); if (e != cudaSuccess) return e;
  double* X = ws.ptr; double* Zw = X + kk; double* sig = Zw + kk; int* rot = reinterpret_cast<int*>(sig + k);
  k_jg_init<<<(unsigned)((kk + 255) / 256), 256, 0, st>>>(W, ldw, k, transpose_in, X, Zw);
  const int n = (k + 1) & ~1, npairs = n >> 1;
  const double tol = sqrt((double)k) * (0.5 * DBL_EPSILON);
  int sweep = 0, total = 0, nl = 1; bool converged = (k < 2);
  while (!converged && sweep < max_sweeps) {
    e = cudaMemsetAsync(rot, 0, sizeof(int), st); if (e != cudaSuccess) return e;
    for (int step = 0; step < n - 1; ++step) k_jg_step<<<npairs, JG_THREADS, 0, st>>>(X, Zw, k, step, tol * tol, rot);
    nl += n - 1;
    int h = 0;
    e = cudaMemcpyAsync(&h, rot, sizeof(int), cudaMemcpyDeviceToHost, st); if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st); if (e != cudaSuccess) return e;       // one read-back per sweep decides whether another sweep is needed
    ++sweep; total += h; converged = (h == 0);
  }
  k_jg_norms<<<(k + 7) / 8, 256, 0, st>>>(X, k, sig);
  k_jg_finish<<<k, 256, 0, st>>>(X, Zw, sig, k, U, ldu, S, Z, ldz, d_info, converged ? sweep : -sweep, total);
  nl += 2;
  if (launches) *launches += nl;
  return cudaGetLastError();
}

}  // namespace

cudaError_t jacobi_svd_square(GemmWorkspace& ws, cudaStream_t st, const double* W, long long ldw, int k, int transpose_in,
                              double* U, long long ldu, double* S, double* Z, long long ldz, int* d_info, int* launches) {
  if (k <= 0) return cudaSuccess;
  // beyond one SM's reach (k > 512) -- and already above k = grid_min(), where the single-CTA global-memory loop is slower --
  // the pairs of a step are spread over the grid
  if (k > grid_min()) return jacobi_svd_grid(ws, st, W, ldw, k, transpose_in, U, ldu, S, Z, ldz, d_info, 60, launches);
  {
    const int sr = (k + JC - 1) / JC, npairs = ((k + 1) & ~1) >> 1;
    const size_t cl_smem = (2 * (size_t)k * sr + 2 * (size_t)npairs * JC * 4 + (size_t)k * JC + k) * sizeof(double);
    static const bool cl_off = getenv("RSVDB_JACOBI_NO_CLUSTER") != nullptr;
    static const bool cl_old = getenv("RSVDB_JACOBI_CL_OLD") != nullptr;       // the barrier.cluster-per-step version, kept for A/B timing
    // measured: the per-step cluster barrier outweighs the bandwidth gain below k ~ 80 (k = 50: 0.99 vs 0.88 ms, k = 100: 1.18 vs 1.40 ms)
    if (!cl_off && k >= 80 && cl_smem + 2 * npairs * 8 <= 220 * 1024 && sr <= 64 && npairs <= 128) {
      const int rplc = (sr + 15) / 16;
#define JCL(R)                                                                                                         \
      { static DevOnce attr;                                                                                           \
        if (!attr.get()) { cudaError_t e = cudaFuncSetAttribute(k_jacobi_cl<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); \
                     if (e == cudaSuccess) e = cudaFuncSetAttribute(k_jacobi_cl2<R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); \
                     if (e == cudaSuccess) e = cudaFuncSetAttribute(k_jacobi_cl2<R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); \
                     if (e != cudaSuccess) return e; attr.set(); }                                                    \
        if (cl_old) k_jacobi_cl<R><<<JC, 1024, cl_smem, st>>>(W, ldw, k, transpose_in, U, ldu, S, Z, ldz, 60, d_info);  \
        else if (npairs <= 64) k_jacobi_cl2<R, 1><<<JC, 1024, cl_smem + 2 * npairs * 8, st>>>(W, ldw, k, transpose_in, U, ldu, S, Z, ldz, 60, d_info); \
        else k_jacobi_cl2<R, 2><<<JC, 1024, cl_smem + 2 * npairs * 8, st>>>(W, ldw, k, transpose_in, U, ldu, S, Z, ldz, 60, d_info); }
      if (rplc <= 1) JCL(1) else if (rplc <= 2) JCL(2) else if (rplc <= 3) JCL(3) else JCL(4)
#undef JCL
      if (launches) ++*launches;
      return cudaGetLastError();
    }
  }
  const size_t smem_need = (2 * (size_t)k * k + k) * sizeof(double);
  const int use_smem = smem_need <= 220 * 1024;
  // neither one SM's nor a 4-CTA cluster's shared memory holds the factor (k > ~116): measured, the grid variant beats a single
  // CTA looping over global memory by 6x at k = 256 and 19x at k = 512 (profiles/r02_jacobi_grid.log)
  if (!use_smem) return jacobi_svd_grid(ws, st, W, ldw, k, transpose_in, U, ldu, S, Z, ldz, d_info, 60, launches);
  double* Xg = nullptr; double* Zg = nullptr;
  const size_t smem = smem_need;
  const int threads = k >= 48 ? 1024 : (k >= 24 ? 512 : 256);
  const int rpl = (k + 15) / 16;
#define JLAUNCH(R)                                                                                                   \
  { static DevOnce attr;                                                                                             \
    if (!attr.get()) { cudaError_t e = cudaFuncSetAttribute(k_jacobi<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); \
                 if (e != cudaSuccess) return e; attr.set(); }                                                      \
    k_jacobi<R><<<1, threads, smem, st>>>(W, ldw, k, transpose_in, U, ldu, S, Z, ldz, Xg, Zg, use_smem, 60, d_info); }
  if (rpl <= 1) JLAUNCH(1) else if (rpl <= 2) JLAUNCH(2) else if (rpl <= 4) JLAUNCH(4) else if (rpl <= 7) JLAUNCH(7) else JLAUNCH(8)
#undef JLAUNCH
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace rsvdb
