// K7: CSR sparse x dense products and the sparse rSVD pipeline (see spmm.cu).
#pragma once
#include <cstdint>
#include "context.cuh"

namespace rsvdb {

// Y_rm (m x l, ROW-major, ld = l) = A (CSR, m x n) * X_rm (n x l, ROW-major, ld = l)
int csr_spmm_rm(rsvdb_ctx* c, int64_t m, const int64_t* rowptr, const int32_t* col, const double* val, const double* X_rm, int l,
                double* Y_rm);
// Explicit transpose of a CSR matrix on the device (stable: entries of a column keep their row order => deterministic sums).
// Outputs are carved from c->io_ws after `io_offset_doubles`; returns pointers through the out arguments.
int csr_transpose(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* col, const double* val,
                  int64_t* rowptrT, int32_t* colT, double* valT);
// rSVD (reference src/rSVD.cpp:72-133) of a CSR matrix held on the device; this rank's row block when c->nranks > 1.
int rsvd_csr_device(rsvdb_ctx* c, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* col, const double* val,
                    const double* Omega, int64_t ldo, int l, int q, int method, double* U, int64_t ldu, double* S, double* V,
                    int64_t ldv, uint64_t seed);

}  // namespace rsvdb
