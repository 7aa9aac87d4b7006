// Host-side entry points of the skinny FP64 GEMM kernels (see gemm_dmma.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace rsvdb {

// Grow-only device scratch buffer (split-K partials, TSQR stacks, ...).
struct GemmWorkspace {
  double* ptr = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) { cudaError_t e = cudaFree(ptr); ptr = nullptr; bytes = 0; if (e != cudaSuccess) return e; }
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

// Y (M x N) = A (M x K) * X (K x N); all column-major.
cudaError_t gemm_an(GemmWorkspace& ws, cudaStream_t st, int nsm, const double* A, long long M, long long K, long long lda,
                    const double* X, long long ldx, int N, double* Y, long long ldy, int* launches);
// A is K x M.  transpose_out == 0: Z (M x N) = A^T * Q;  transpose_out == 1: Z (N x M) = Q^T * A.
cudaError_t gemm_at(GemmWorkspace& ws, cudaStream_t st, int nsm, const double* A, long long K, long long M, long long lda,
                    const double* Q, long long ldq, int N, double* Z, long long ldz, int transpose_out, int* launches);
// C (m x n) = alpha * op(A) * op(B) + beta * C on CUDA cores (small / unaligned operands).
cudaError_t gemm_generic(cudaStream_t st, int ta, int tb, int m, int n, int k, double alpha, const double* A, long long lda,
                         const double* B, long long ldb, double beta, double* C, long long ldc);

// Performance notes (process-wide counters, not errors).  split_product_count: gemm_an / gemm_at calls whose A had an odd
// leading dimension or an 8-byte-aligned base and therefore ran on the two-tensor-map variant of the DMMA kernels (same
// speed class).  generic_fallback_count: calls that ran on the CUDA-core kernel (only a 1-column / 1-row A that TMA cannot describe).
long long generic_fallback_count();
long long split_product_count();

}  // namespace rsvdb
