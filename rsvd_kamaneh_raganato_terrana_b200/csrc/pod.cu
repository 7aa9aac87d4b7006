// POD wrappers (reference POD/ParametricDiffusion1D/src/POD.cpp).  The snapshot matrix S (Nh x ns) stays on the device;
// the correlation matrix S^T S (or S S^T, or the energy / weight variants) is formed with the DMMA GEMMs, handed to the
// SVD back-end the caller selects (perform_SVD, POD.cpp:42-114), and the modes are recovered with one more skinny GEMM.
// Reference quirks that are kept because callers see them: sigma holds the singular values of the CORRELATION matrix
// (= sigma(S)^2) and the modes are divided by it (POD.cpp:164-166); sigma is returned at the back-end's full length.
#include "pod.cuh"

#include "../../include/rsvdb.h"

#include <algorithm>
#include <cmath>
#include <vector>

#include "pipeline.cuh"

namespace rsvdb {
namespace {

// out(:, i) = in(:, i) / sigma[i]                        (POD.cpp:165  W.col(i) = S*V.col(i)/sigma(i))
__global__ void k_div_cols(double* __restrict__ W, long long ld, long long rows, int cols, const double* __restrict__ sigma) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) W[(size_t)j * ld + i] = W[(size_t)j * ld + i] / sigma[j];
}
// out(:, j) = in(:, j) * f(s[j]),  f = sqrt (mode 0) or 1/sqrt (mode 1)      (symmetric square roots)
__global__ void k_scale_cols_fn(const double* __restrict__ in, long long ldi, double* __restrict__ out, long long ldo, int rows,
                                int cols, const double* __restrict__ s, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int j = blockIdx.y; j < cols; j += gridDim.y) {
    const double f = mode == 0 ? sqrt(s[j]) : 1.0 / sqrt(s[j]);
    out[(size_t)j * ldo + i] = in[(size_t)j * ldi + i] * f;
  }
}
// First r columns of the layout SVD<Power> gives V: an identity-initialised b x b matrix whose ROW j is the j-th right
// singular vector (include/SVD_class.hpp:83,214).  Vc holds the vectors as columns (b x k).
__global__ void k_power_vref(const double* __restrict__ Vc, long long ldv, int b, int k, int r, double* __restrict__ out, long long ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;   // row of the reference layout
  if (j >= b) return;
  for (int i = blockIdx.y; i < r; i += gridDim.y) out[(size_t)i * ldo + j] = (j < k) ? Vc[(size_t)j * ldv + i] : (j == i ? 1.0 : 0.0);
}

struct Arena {   // bump allocator over pod_ws; an undersized plan is reported (overflow) instead of running past the buffer
  double* base; size_t off = 0, cap; bool overflow = false;
  double* take(size_t d) {
    const size_t padded = (d + 31) & ~size_t(31);
    if (off + padded > cap) { overflow = true; return base; }
    double* p = base + off; off += padded; return p;
  }
};

struct SvdOut {
  double* U = nullptr; int64_t ldu = 0, ucols = 0;     // a x ucols
  double* sigma = nullptr; int64_t slen = 0;
  double* Vr = nullptr; int64_t ldvr = 0;              // first r columns of V in the reference's layout (b x r)
};

size_t svd_scratch_doubles(int64_t a, int64_t b, int r, int svd_type) {
  const int64_t k = std::min(a, b);
  switch (svd_type) {
    case 0: return (size_t)a * b + (size_t)a * a + (size_t)b * r + k + (size_t)b * r + 256;
    case 1: case 2: return (size_t)a * k + (size_t)b * k + k + 256;
    default: return (size_t)b * r + (size_t)a * r + (size_t)b * r + r + (size_t)b * r + 256;
  }
}

// perform_SVD (POD.cpp:42-114) on a device matrix M (a x b).  svd_type: 0 Power, 1 Jacobi, 2 ParallelJacobi,
// 3/4/5 rSVD with the Power / Jacobi / ParallelJacobi back-end and l = r.
int perform_svd(rsvdb_ctx* c, Arena& ar, const double* M, int64_t a, int64_t b, int64_t ldm, int r, int svd_type, uint64_t seed,
                const double* Omega, int64_t ldo, SvdOut* o) {
  const int64_t k = std::min(a, b);
  if (r <= 0) return fail(c, -1, "POD: r must be positive");
  switch (svd_type) {
    case 0: {                                            // SVD<Power>(A, r)            include/SVD_class.hpp:184-219
      if (r > k) return fail(c, -1, "POD: r larger than min(rows, cols) of the matrix handed to SVD<Power>");
      if (a > 16384) return fail(c, -6, "POD: SVD<Power> returns a rows x rows U; more than 16384 rows is not supported");
      double* Mt = ar.take((size_t)b * a);
      o->U = ar.take((size_t)a * a); o->ldu = a; o->ucols = a;
      o->sigma = ar.take((size_t)k); o->slen = k;
      double* Vc = ar.take((size_t)b * r);
      o->Vr = ar.take((size_t)b * r); o->ldvr = b;
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      RSVDB_TRY(transpose2d(c, M, ldm, Mt, b, a, b));
      int found = 0;
      RSVDB_TRY(small_svd_power_t(c, Mt, b, a, b, r, seed, o->U, a, (int)a, o->sigma, Vc, b, &found));
      if (found < r) return fail(c, -5, "POD: SVD<Power> met a singular value below 1e-12 before r were found (the reference then indexes past the end)");
      k_power_vref<<<dim3((unsigned)((b + 255) / 256), (unsigned)std::min(r, 64)), 256, 0, c->stream>>>(Vc, b, (int)b, r, r, o->Vr, b);
      RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
      return 0;
    }
    case 1: case 2: {                                    // SVD<Jacobi | ParallelJacobi>(A)
      if (r > k) return fail(c, -1, "POD: r larger than min(rows, cols) of the matrix handed to the SVD");
      o->U = ar.take((size_t)a * k); o->ldu = a; o->ucols = k;
      o->sigma = ar.take((size_t)k); o->slen = k;
      double* V = ar.take((size_t)b * k);
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      RSVDB_TRY(small_svd_jacobi(c, M, ldm, nullptr, 0, a, b, o->U, a, o->sigma, V, b));
      o->Vr = V; o->ldvr = b;
      return 0;
    }
    case 3: case 4: case 5: {                            // rSVD(A, U, sigma, V, r, method)   src/rSVD.cpp:72-133
      const int method = svd_type == 3 ? 1 : (svd_type == 4 ? 0 : 2);
      if (r > b) return fail(c, -1, "POD: r larger than the column count of the matrix handed to rSVD");
      double* Om = nullptr; int64_t ldom = b;
      if (Omega) { Om = const_cast<double*>(Omega); ldom = ldo; }
      else Om = ar.take((size_t)b * r);
      o->U = ar.take((size_t)a * r); o->ldu = a; o->ucols = r;
      o->sigma = ar.take((size_t)r); o->slen = r;
      double* V = ar.take((size_t)b * r);
      double* Vr_pow = method == 1 ? ar.take((size_t)b * r) : nullptr;
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      if (!Omega) RSVDB_TRY(rsvdb_generate_omega_dev(c, b, r, seed, Om, b));
      RSVDB_TRY(rsvd_device(c, M, a, b, ldm, Om, ldom, r, 2, method, o->U, a, o->sigma, V, b, seed));
      if (method == 1) {                                 // V is b x b with the vectors in rows
        o->Vr = Vr_pow; o->ldvr = b;
        k_power_vref<<<dim3((unsigned)((b + 255) / 256), (unsigned)std::min(r, 64)), 256, 0, c->stream>>>(V, b, (int)b, r, r, o->Vr, b);
        RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
      } else { o->Vr = V; o->ldvr = b; }
      return 0;
    }
    default:
      return fail(c, -1, "The svd_type should be in [0,5]. Check 'svd_type' in the parameter file.");   // POD.cpp:87-91
  }
}

int gg(rsvdb_ctx* c, int ta, int tb, int64_t m, int64_t n, int64_t k, const double* A, int64_t lda, const double* B, int64_t ldb,
       double* C, int64_t ldc) {
  RSVDB_CUDA(c, gemm_generic(c->stream, ta, tb, (int)m, (int)n, (int)k, 1.0, A, lda, B, ldb, 0.0, C, ldc));
  ++c->launches;
  return 0;
}

// Symmetric positive definite X (n x n): Xs = X^{1/2} (SelfAdjointEigenSolver::operatorSqrt in the reference, POD.cpp:272-273)
// and optionally Xis = X^{-1/2}; through the Jacobi SVD X = U diag(s) V^T (U = V for an SPD matrix).
int spd_sqrt(rsvdb_ctx* c, Arena& ar, const double* X, int64_t n, int64_t ldx, double* Xs, double* Xis) {
  if (n > 512) return fail(c, -6, "POD: matrix square root of an operator larger than 512 x 512 is not supported");
  double* U = ar.take((size_t)n * n); double* V = ar.take((size_t)n * n); double* s = ar.take((size_t)n); double* T = ar.take((size_t)n * n);
  if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
  RSVDB_TRY(small_svd_jacobi(c, X, ldx, nullptr, 0, n, n, U, n, s, V, n));
  const dim3 g((unsigned)((n + 255) / 256), (unsigned)std::min<int64_t>(n, 64));
  k_scale_cols_fn<<<g, 256, 0, c->stream>>>(V, n, T, n, (int)n, (int)n, s, 0);
  RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  RSVDB_TRY(gg(c, 0, 1, n, n, n, T, n, V, n, Xs, n));
  if (Xis) {
    k_scale_cols_fn<<<g, 256, 0, c->stream>>>(V, n, T, n, (int)n, (int)n, s, 1);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
    RSVDB_TRY(gg(c, 0, 1, n, n, n, T, n, V, n, Xis, n));
  }
  return 0;
}

}  // namespace

bool pod_shape(int variant, int64_t Nh, int64_t ns, int r, int svd_type, PodShape* out) {
  if (svd_type < 0 || svd_type > 5) return false;
  if (variant == POD_NAIVE) {                            // W = U of perform_SVD(S): POD.cpp:116-134
    const int64_t k = std::min(Nh, ns);
    out->w_cols_full = svd_type == 0 ? Nh : (svd_type <= 2 ? k : r);
    out->sigma_len = svd_type <= 2 ? k : r;
    return true;
  }
  const int64_t d = ns <= Nh ? ns : Nh;                  // the correlation matrix is d x d
  out->sigma_len = svd_type <= 2 ? d : r;
  if (ns <= Nh) out->w_cols_full = r;                    // W stays Nh x r (POD.cpp:149,164-166)
  else if (variant == POD_STANDARD) out->w_cols_full = svd_type <= 2 ? d : r;   // W = U (POD.cpp:190)
  else out->w_cols_full = r;   // U is Nh x r (POD.cpp:293); with svd_type 0-2 the reference's solve loop (:300-302) runs past it
  return true;
}

int pod_device(rsvdb_ctx* c, int variant, const double* S, int64_t Nh, int64_t ns, int64_t lds, const double* Xh, int64_t ldx,
               const double* D, int64_t ldd, int r, double tol, int svd_type, uint64_t seed, const double* Omega, int64_t ldo,
               double* W, int64_t ldw, double* sigma, int* N) {
  PodShape shp;
  if (!pod_shape(variant, Nh, ns, r, svd_type, &shp))
    return fail(c, -1, "The svd_type should be in [0,5]. Check 'svd_type' in the parameter file.");
  if (variant < 0 || variant > 3 || Nh <= 0 || ns <= 0 || r <= 0) return fail(c, -1, "POD: bad argument");
  if ((variant >= POD_ENERGY && !Xh) || (variant == POD_WEIGHT && !D)) return fail(c, -1, "POD: the energy / weight variants need Xh (and D)");
  const int64_t d = ns <= Nh ? ns : Nh;
  // workspace
  size_t need = 0;
  if (variant == POD_NAIVE) need = svd_scratch_doubles(Nh, ns, r, svd_type);
  else {
    need = svd_scratch_doubles(d, d, r, svd_type) + (size_t)d * d + 64;
    if (ns > Nh) need += (size_t)ns * Nh + 64;                                        // S^T
    if (variant >= POD_ENERGY) {
      if (ns <= Nh) {   // T = S^T Xh; weight: D^{1/2}, S D^{1/2} and the square root's scratch (U, V, T, s)
        need += (size_t)ns * Nh + 256;
        if (variant == POD_WEIGHT) need += 4 * (size_t)ns * ns + (size_t)Nh * ns + (size_t)ns + 1024;
      } else {          // Xh^{1/2}, Xh^{-1/2}, the square root's scratch, T1, T2, T3
        need += 6 * (size_t)Nh * Nh + 2 * (size_t)Nh * ns + (size_t)Nh + 1024;
      }
    }
  }
  need += 4096;
  RSVDB_CUDA(c, c->pod_ws.reserve(need * sizeof(double)));
  Arena ar{c->pod_ws.ptr, 0, need};
  SvdOut o;
  PhaseTimer pt(c, PH_OTHER);
  int nl = 0;

  if (variant == POD_NAIVE) {
    RSVDB_TRY(perform_svd(c, ar, S, Nh, ns, lds, r, svd_type, seed, Omega, ldo, &o));
    RSVDB_TRY(copy2d(c, o.U, o.ldu, W, ldw, Nh, (int)o.ucols));
    RSVDB_CUDA(c, cudaMemcpyAsync(sigma, o.sigma, (size_t)o.slen * 8, cudaMemcpyDeviceToDevice, c->stream));
    *N = (int)o.ucols;
    return 0;
  }

  const double* Smodes = S; int64_t ld_modes = lds;      // the matrix the modes are recovered from (S or S * D^{1/2})
  double* C = ar.take((size_t)d * d);
  if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
  double* Xis = nullptr;                                 // Xh^{-1/2} for the ns > Nh energy / weight branches
  if (ns <= Nh) {
    if (variant == POD_STANDARD) {                       // C = S^T S                         POD.cpp:152
      RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, S, Nh, ns, lds, S, lds, (int)ns, C, ns, 0, &nl));
    } else {
      if (variant == POD_WEIGHT) {                       // Stilde = S * D^{1/2}              POD.cpp:363-370
        double* Ds = ar.take((size_t)ns * ns); double* St = ar.take((size_t)Nh * ns);
        if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
        RSVDB_TRY(spd_sqrt(c, ar, D, ns, ldd, Ds, nullptr));
        RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, S, Nh, ns, lds, Ds, ns, (int)ns, St, Nh, &nl));
        Smodes = St; ld_modes = Nh;
      }
      // Ctilde = (S^T Xh) S                                                          POD.cpp:250, :373
      double* T = ar.take((size_t)ns * Nh);
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, Smodes, Nh, ns, ld_modes, Xh, ldx, (int)Nh, T, ns, 0, &nl));
      RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, T, ns, Nh, ns, Smodes, ld_modes, (int)ns, C, ns, &nl));
    }
  } else {
    if (variant == POD_STANDARD) {                       // K = S S^T                         POD.cpp:171
      double* St = ar.take((size_t)ns * Nh);
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      RSVDB_TRY(transpose2d(c, S, lds, St, ns, Nh, ns));
      RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, St, ns, Nh, ns, St, ns, (int)Nh, C, Nh, 0, &nl));
    } else {                                             // Ktilde = Xs S [D] S^T Xs          POD.cpp:272-280, :402-409
      double* Xs = ar.take((size_t)Nh * Nh); Xis = ar.take((size_t)Nh * Nh);
      RSVDB_TRY(spd_sqrt(c, ar, Xh, Nh, ldx, Xs, Xis));
      double* T1 = ar.take((size_t)Nh * ns); double* T2 = ar.take((size_t)Nh * ns); double* T3 = ar.take((size_t)Nh * Nh);
      if (ar.overflow) return fail(c, -4, "POD: workspace plan too small");
      RSVDB_TRY(gg(c, 0, 0, Nh, ns, Nh, Xs, Nh, S, lds, T1, Nh));
      const double* L = T1;
      if (variant == POD_WEIGHT) { RSVDB_TRY(gg(c, 0, 0, Nh, ns, ns, T1, Nh, D, ldd, T2, Nh)); L = T2; }
      RSVDB_TRY(gg(c, 0, 1, Nh, Nh, ns, L, Nh, S, lds, T3, Nh));
      RSVDB_TRY(gg(c, 0, 0, Nh, Nh, Nh, T3, Nh, Xs, Nh, C, Nh));
    }
  }
  c->launches += nl; nl = 0;

  RSVDB_TRY(perform_svd(c, ar, C, d, d, d, r, svd_type, seed, Omega, ldo, &o));
  if (o.slen < r) return fail(c, -1, "POD: the SVD back-end returned fewer than r singular values");

  if (ns <= Nh) {                                        // W.col(i) = S V.col(i) / sigma(i)   POD.cpp:164-166
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, Smodes, Nh, ns, ld_modes, o.Vr, o.ldvr, r, W, ldw, &nl));
    c->launches += nl;
    k_div_cols<<<dim3((unsigned)((Nh + 255) / 256), (unsigned)std::min(r, 32)), 256, 0, c->stream>>>(W, ldw, Nh, r, o.sigma);
    RSVDB_CUDA(c, cudaGetLastError()); ++c->launches;
  } else if (variant == POD_STANDARD) {                  // W = U                              POD.cpp:190
    RSVDB_TRY(copy2d(c, o.U, o.ldu, W, ldw, Nh, (int)o.ucols));
  } else {                                               // Xh^{1/2} U = Utilde (CG to 1e-12 in the reference, POD.cpp:296-304)
    RSVDB_TRY(gg(c, 0, 0, Nh, std::min<int64_t>(o.ucols, r), Nh, Xis, Nh, o.U, o.ldu, W, ldw));
  }
  RSVDB_CUDA(c, cudaMemcpyAsync(sigma, o.sigma, (size_t)o.slen * 8, cudaMemcpyDeviceToDevice, c->stream));

  // energy criterion on the first r singular values (POD.cpp:203-219): smallest N with sum_{i<N} s_i^2 / sum_{i<r} s_i^2 >= 1 - tol^2
  std::vector<double> hs((size_t)r);
  RSVDB_CUDA(c, cudaMemcpyAsync(hs.data(), o.sigma, (size_t)r * 8, cudaMemcpyDeviceToHost, c->stream));
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  double den = 0.0;
  for (int i = 0; i < r; ++i) den += std::pow(hs[(size_t)i], 2);
  int n_keep = 0; double I = 0.0, num = 0.0;
  while (I < (1 - std::pow(tol, 2)) && n_keep < r) { num += std::pow(hs[(size_t)n_keep], 2); I = num / den; ++n_keep; }
  *N = n_keep;
  return 0;
}

}  // namespace rsvdb
