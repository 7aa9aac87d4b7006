// Small dense SVD by one-sided (Hestenes) Jacobi (see jacobi.cu).
#pragma once
#include <cuda_runtime.h>
#include "gemm_dmma.cuh"

namespace rsvdb {

// W is k x k (ldw) on the device; transpose_in != 0 reads W^T instead.  Computes W = U diag(S) Z^T with S descending,
// U (k x k, ldu), Z (k x k, ldz).  info[0] = sweeps used (negative: not converged), info[1] = rotations applied.
cudaError_t jacobi_svd_square(GemmWorkspace& ws, cudaStream_t st, const double* W, long long ldw, int k, int transpose_in,
                              double* U, long long ldu, double* S, double* Z, long long ldz, int* d_info, int* launches);

}  // namespace rsvdb
