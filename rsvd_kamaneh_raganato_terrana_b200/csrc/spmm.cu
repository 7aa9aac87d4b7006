// K7: FP64 CSR SpMM for sparse MatrixMarket inputs.
//
// The reference has no sparse path: every .mtx is densified before rSVD (tests/rSVD_test.cpp:54-57,
// `MatrixXd(sparseMatrix)`), which is impossible for BASELINE.json's 1M x 1M, ~10 nnz/row variant (8 TB dense).  Here
// the two products of the range finder, Y = A X and Z = A^T Q (src/rSVD.cpp:59,63,66,89), run on the CSR directly:
//   * dense operands are held ROW-major inside the sparse pipeline, so that the gather of row `col[k]` of X is one
//     contiguous, vectorised (16-byte) read of l doubles -- the access that dominates an SpMM with random columns;
//   * one warp per matrix row; the row's (value, column) pairs are loaded coalesced, broadcast with shuffles, and the
//     gathers of four non-zeros are issued back to back before their FMAs (memory-level parallelism);
//   * A^T Q uses an explicitly transposed CSR built once per rSVD with a STABLE radix sort by column (CUB), so the
//     transposed product is the same gather kernel: no atomics, bit-reproducible sums.
// This is an HBM-bound kernel: bench/tools report GB/s against the measured copy bandwidth, both for the compulsory
// bytes (12 nnz + 8 (m+1) + 8 l (m+n)) and including the gathers (8 l nnz), cf. SURVEY.md 8d.
#include "spmm.cuh"

#include <algorithm>
#include <cub/cub.cuh>

#include "comm.cuh"
#include "pipeline.cuh"

namespace rsvdb {

namespace {

constexpr int SPMM_WARPS = 8;

// VEC2: l even and operands 16-byte aligned -> double2 accesses.  LCH = ceil(l / 64) column chunks per lane (VEC2) or
// ceil(l / 32) (scalar); accumulators stay in registers.
template <int LCH, bool VEC2>
__global__ void __launch_bounds__(SPMM_WARPS * 32)
k_csr_spmm_rm(long long m, const long long* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
              const double* __restrict__ X, int l, double* __restrict__ Y) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * SPMM_WARPS;
  for (long long i = warp; i < m; i += nwarps) {
    const long long p0 = rowptr[i], p1 = rowptr[i + 1];
    double acc[LCH][2];
#pragma unroll
    for (int h = 0; h < LCH; ++h) acc[h][0] = acc[h][1] = 0.0;
    for (long long p = p0; p < p1; p += 32) {
      const int cnt = (int)min((long long)32, p1 - p);
      double v = 0.0; int cidx = 0;
      if (lane < cnt) { v = val[p + lane]; cidx = col[p + lane]; }
      for (int j0 = 0; j0 < cnt; j0 += 4) {
        double vv[4]; const double* xr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, cnt - 1);
          vv[u] = (j0 + u < cnt) ? __shfl_sync(0xffffffffu, v, j) : (__shfl_sync(0xffffffffu, v, j), 0.0);
          xr[u] = X + (size_t)__shfl_sync(0xffffffffu, cidx, j) * l;
        }
        if (VEC2) {
          double2 x[4][LCH];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < LCH; ++h) {
              const int cc = 2 * lane + 64 * h;
              x[u][h] = (cc < l) ? *reinterpret_cast<const double2*>(xr[u] + cc) : make_double2(0.0, 0.0);
            }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < LCH; ++h) { acc[h][0] = fma(vv[u], x[u][h].x, acc[h][0]); acc[h][1] = fma(vv[u], x[u][h].y, acc[h][1]); }
        } else {
          double x[4][LCH];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < LCH; ++h) { const int cc = lane + 32 * h; x[u][h] = (cc < l) ? xr[u][cc] : 0.0; }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int h = 0; h < LCH; ++h) acc[h][0] = fma(vv[u], x[u][h], acc[h][0]);
        }
      }
    }
    double* yr = Y + (size_t)i * l;
#pragma unroll
    for (int h = 0; h < LCH; ++h) {
      if (VEC2) { const int cc = 2 * lane + 64 * h; if (cc < l) *reinterpret_cast<double2*>(yr + cc) = make_double2(acc[h][0], acc[h][1]); }
      else { const int cc = lane + 32 * h; if (cc < l) yr[cc] = acc[h][0]; }
    }
  }
}

__global__ void k_expand_rows(long long m, const long long* __restrict__ rowptr, int* __restrict__ rowidx) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < m; i += nwarps)
    for (long long p = rowptr[i] + lane; p < rowptr[i + 1]; p += 32) rowidx[p] = (int)i;
}
__global__ void k_iota(long long n, int* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)i;
}
// rowptrT[c] = first position in the column-sorted key array whose key is >= c (binary search: no atomics)
__global__ void k_lower_bounds(long long n, long long nnz, const int* __restrict__ sorted_keys, long long* __restrict__ rowptrT) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n) return;
  long long lo = 0, hi = nnz;
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (sorted_keys[mid] < (int)c) lo = mid + 1; else hi = mid; }
  rowptrT[c] = lo;
}
__global__ void k_permute(long long nnz, const int* __restrict__ perm, const int* __restrict__ rowidx, const double* __restrict__ val,
                          int* __restrict__ colT, double* __restrict__ valT) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) { const int src = perm[i]; colT[i] = rowidx[src]; valT[i] = val[src]; }
}

}  // namespace

int csr_spmm_rm(rsvdb_ctx* c, int64_t m, const int64_t* rowptr, const int32_t* col, const double* val, const double* X, int l, double* Y) {
  if (m <= 0 || l <= 0) return 0;
  if (l > 512) return fail(c, -6, "csr_spmm: l > 512 is not supported");
  const bool vec2 = (l % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
  const int grid = (int)std::min<int64_t>((m + SPMM_WARPS - 1) / SPMM_WARPS, (int64_t)c->nsm * 16);
  const long long* rp = reinterpret_cast<const long long*>(rowptr);
#define SPMM_LAUNCH(LCHV, V2) k_csr_spmm_rm<LCHV, V2><<<grid, SPMM_WARPS * 32, 0, c->stream>>>(m, rp, col, val, X, l, Y)
  if (vec2) {
    const int lch = (l + 63) / 64;
    switch (lch) { case 1: SPMM_LAUNCH(1, true); break; case 2: SPMM_LAUNCH(2, true); break; case 3: SPMM_LAUNCH(3, true); break;
                   case 4: SPMM_LAUNCH(4, true); break; default: SPMM_LAUNCH(8, true); break; }
  } else {
    const int lch = (l + 31) / 32;
    switch (lch) { case 1: SPMM_LAUNCH(1, false); break; case 2: SPMM_LAUNCH(2, false); break; case 3: SPMM_LAUNCH(3, false); break;
                   case 4: SPMM_LAUNCH(4, false); break; case 5: case 6: case 7: case 8: SPMM_LAUNCH(8, false); break; default: SPMM_LAUNCH(16, false); break; }
  }
#undef SPMM_LAUNCH
  RSVDB_CUDA(c, cudaGetLastError());
  ++c->launches;
  return 0;
}

int csr_transpose(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* col, const double* val,
                  int64_t* rowptrT, int32_t* colT, double* valT) {
  if (nnz >= (1LL << 31)) return fail(c, -6, "csr_transpose: nnz >= 2^31 per rank is not supported");
  cudaStream_t st = c->stream;
  if (nnz == 0) { RSVDB_CUDA(c, cudaMemsetAsync(rowptrT, 0, (size_t)(n + 1) * 8, st)); return 0; }
  // scratch in svd_ws: rowidx, iota, sorted keys, perm (4 x nnz int32) + CUB temp
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, (int)nnz, 0, 32, st);
  const size_t ints = ((size_t)nnz + 3) & ~size_t(3);
  RSVDB_CUDA(c, c->svd_ws.reserve(4 * ints * sizeof(int) + temp_bytes + 256));
  int* rowidx = reinterpret_cast<int*>(c->svd_ws.ptr); int* iota = rowidx + ints; int* keys = iota + ints; int* perm = keys + ints;
  void* temp = perm + ints;
  const int T = 256;
  k_expand_rows<<<(int)std::min<int64_t>((m * 32 + T - 1) / T, 65535), T, 0, st>>>(m, reinterpret_cast<const long long*>(rowptr), rowidx);
  k_iota<<<(int)((nnz + T - 1) / T), T, 0, st>>>(nnz, iota);
  int bits = 1; while (bits < 32 && (1LL << bits) < n) ++bits;
  RSVDB_CUDA(c, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col, keys, iota, perm, (int)nnz, 0, bits, st));   // LSD radix sort: stable
  k_lower_bounds<<<(int)((n + 1 + T - 1) / T), T, 0, st>>>(n, nnz, keys, reinterpret_cast<long long*>(rowptrT));
  k_permute<<<(int)((nnz + T - 1) / T), T, 0, st>>>(nnz, perm, rowidx, val, colT, valT);
  RSVDB_CUDA(c, cudaGetLastError());
  c->launches += 5;
  return 0;
}

int rsvd_csr_device(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* col, const double* val,
                    const double* Omega, int64_t ldo, int l, int q, int method, double* U, int64_t ldu, double* S, double* V,
                    int64_t ldv, uint64_t seed) {
  if (method != 0 && method != 1 && method != 2) return fail(c, -1, "Unsupported SVD method");
  if (l <= 0 || n <= 0 || m < 0 || q < 0) return fail(c, -1, "rSVD (CSR): bad shape");
  const int64_t k = std::min<int64_t>(l, n);
  const int64_t big = std::max(m, n);
  // tmp_ws: [Bt n x l][Q m x l][Ut l x l][R1 big x l (row-major staging)][R2 big x l][transposed CSR]
  // even leading dimensions / offsets: the dense products on Q and Bt stay on the TMA + DMMA path for odd m, n, l
  const int64_t ldb = even_ld(n), ldq = even_ld(m), ldut = even_ld(l);
  const size_t d_bt = (size_t)ldb * l, d_q = (size_t)ldq * l, d_ut = (size_t)ldut * l + 64, d_r = (((size_t)big * l + 1) & ~size_t(1)) + 64;
  const size_t d_rowptrT = (size_t)n + 2, d_colT = ((size_t)nnz + 1) / 2 + 2, d_valT = (size_t)nnz + 2;
  RSVDB_CUDA(c, c->tmp_ws.reserve((d_bt + d_q + d_ut + 2 * d_r + d_rowptrT + d_colT + d_valT) * sizeof(double)));
  double* Bt = c->tmp_ws.ptr; double* Q = Bt + d_bt; double* Ut = Q + d_q; double* R1 = Ut + d_ut; double* R2 = R1 + d_r;
  int64_t* rowptrT = reinterpret_cast<int64_t*>(R2 + d_r); int32_t* colT = reinterpret_cast<int32_t*>(rowptrT + d_rowptrT);
  double* valT = reinterpret_cast<double*>(colT) + d_colT;
  {
    PhaseTimer pt(c, PH_OTHER);
    RSVDB_TRY(csr_transpose(c, m, n, nnz, rowptr, col, val, rowptrT, colT, valT));
  }
  auto spmm_a = [&](const double* Xcm, int64_t ldx, double* Ycm, int64_t ldy) -> int {   // Y (m x l) = A * X (n x l), column-major in/out
    PhaseTimer pt(c, PH_GEMM_AN);
    RSVDB_TRY(transpose2d(c, Xcm, ldx, R1, l, n, l));               // X_rm (n x l row-major) = "l x n column-major"
    RSVDB_TRY(csr_spmm_rm(c, m, rowptr, col, val, R1, l, R2));
    RSVDB_TRY(transpose2d(c, R2, l, Ycm, ldy, l, m));                // back to column-major m x l
    return 0;
  };
  auto spmm_at = [&](const double* Qcm, int64_t ldq, double* Zcm, int64_t ldz) -> int {  // Z (n x l) = A^T * Q (m x l)
    {
      PhaseTimer pt(c, PH_GEMM_AT);
      RSVDB_TRY(transpose2d(c, Qcm, ldq, R1, l, m, l));
      RSVDB_TRY(csr_spmm_rm(c, n, rowptrT, colT, valT, R1, l, R2));
      RSVDB_TRY(transpose2d(c, R2, l, Zcm, ldz, l, n));
    }
    if (c->nranks > 1) { PhaseTimer pc(c, PH_COMM); RSVDB_TRY(comm_allreduce_sum(c, Zcm, (size_t)ldz * l)); }
    return 0;
  };
  RSVDB_TRY(spmm_a(Omega, ldo, Q, ldq));                              // Y = A * Omega                  src/rSVD.cpp:59
  c->chol_failed = false;
  RSVDB_TRY(orthonormalize(c, Q, m, l, ldq, true, nullptr));          // :60-61
  for (int it = 0; it < q; ++it) {
    RSVDB_TRY(spmm_at(Q, ldq, Bt, ldb));                              // Y = A^T * Q                    :63
    RSVDB_TRY(orthonormalize(c, Bt, n, l, ldb, false, nullptr));      // :64-65
    RSVDB_TRY(spmm_a(Bt, ldb, Q, ldq));                               // Y = A * Q                      :66
    RSVDB_TRY(orthonormalize(c, Q, m, l, ldq, true, nullptr));        // :67-68
  }
  RSVDB_TRY(spmm_at(Q, ldq, Bt, ldb));                                // B^T = A^T Q                    :89
  if (method == 1) { RSVDB_TRY(small_svd_power_t(c, Bt, ldb, l, n, 0, seed, Ut, ldut, l, S, V, ldv, nullptr)); }
  else { RSVDB_TRY(small_svd_jacobi(c, nullptr, 0, Bt, ldb, l, n, Ut, ldut, S, V, ldv)); }
  {
    PhaseTimer pt(c, PH_OTHER);
    int nl = 0;
    RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, Q, m, l, ldq, Ut, ldut, (int)k, U, ldu, &nl));   // U = Q * Utilde  :128
    c->launches += nl;
  }
  return 0;
}

}  // namespace rsvdb
