// Engine context behind the C ABI (include/rsvdb.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "gemm_dmma.cuh"
#include "host_stage.cuh"

enum RsvdbPhase { PH_GEMM_AN = 0, PH_GEMM_AT = 1, PH_QR = 2, PH_SMALL_SVD = 3, PH_COMM = 4, PH_OTHER = 5, PH_COPY = 6, PH_COUNT = 7 };

struct rsvdb_ctx {
  int device = 0;
  int nsm = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;   // overlaps the explicit-factor formation of upper TSQR levels with the factor chain
  cudaEvent_t side_ev[20] = {};          // [0..15] per tree level, [16] "side work done"
  rsvdb::GemmWorkspace gemm_ws;     // split-K partial tiles
  rsvdb::GemmWorkspace qr_ws;       // TSQR tree of the local panel
  rsvdb::GemmWorkspace qr2_ws;      // TSQR tree of the all-gathered R stack (multi-GPU)
  rsvdb::GemmWorkspace tmp_ws;      // pipeline intermediates (Q, Z, B^T, small factors)
  rsvdb::GemmWorkspace svd_ws;      // small-SVD scratch
  rsvdb::GemmWorkspace io_ws;       // device copies for the *_host entry points
  rsvdb::GemmWorkspace wide_ws;     // wide-panel QR (l > 100): R, projection coefficients, product buffer
  rsvdb::GemmWorkspace pca_ws;      // column statistics, implicit-centring operands
  rsvdb::GemmWorkspace pod_ws;      // POD: correlation matrix, SVD factors of it
  rsvdb::GemmWorkspace chol_ws;     // CholeskyQR2 fast path: the intermediate panel, Gram matrices, triangular factors
  double* chol_host = nullptr;      // pinned: the guard's verdict (pivot breakdown, ||Q1^T Q1 - I||_F^2)
  int qr_policy = -1;               // -1: environment default (RSVDB_CHOLQR), 0: guarded CholeskyQR2 then Householder, 1: Householder only
  bool chol_failed = false;         // the guard refused a sketch of the factorisation in flight: stay on Householder for the rest of it
  bool in_rsvd = false;
  bool wide_fast = false;           // qr_wide was entered from the pipeline's orthonormalize: its blocks may take the fast path
  int64_t qr_fast = 0, qr_householder = 0;   // sketches orthonormalised by either path (rsvdb_qr_path_counts)
  int64_t launches = 0;
  std::string err;
  // multi-GPU (row-sharded A); comm is an ncclComm_t resolved at run time (comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // optional per-phase device timing (CUDA events on the stream)
  bool profiling = false;
  struct Span { int phase; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> event_pool;
  const int* d_svd_info = nullptr;   // device {sweeps, rotations} of the last Jacobi SVD
  rsvdb::HostStager* stager = nullptr;   // pinned ring + worker threads for uploads from pageable host memory (lazy)
};

namespace rsvdb {
// host -> device copy of a column-major block on `st`: pageable sources of at least 32 MB go through the pinned staging ring
inline cudaError_t upload_block(rsvdb_ctx* c, cudaStream_t st, double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if ((size_t)rows * (size_t)cols * sizeof(double) >= (32u << 20) && HostStager::pageable(src)) {
    if (!c->stager) c->stager = new HostStager();
    return c->stager->upload(st, dst, ldd, src, lds, rows, cols);
  }
  if (ldd == rows && lds == rows) return cudaMemcpyAsync(dst, src, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, st);
  return cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyHostToDevice, st);
}
inline int fail(rsvdb_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
inline int cuda_fail(rsvdb_ctx* c, cudaError_t e, const char* where) {
  if (c) c->err = std::string(where) + ": " + cudaGetErrorString(e);
  return -2;
}
// RAII phase marker: records an event pair around a pipeline phase when profiling is on.
struct PhaseTimer {
  rsvdb_ctx* c; int idx = -1;
  PhaseTimer(rsvdb_ctx* ctx, int phase) : c(ctx) {
    if (!c->profiling) return;
    rsvdb_ctx::Span s; s.phase = phase;
    auto get = [&]() { cudaEvent_t e; if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    s.a = get(); s.b = get();
    cudaEventRecord(s.a, c->stream);
    c->spans.push_back(s); idx = (int)c->spans.size() - 1;
  }
  ~PhaseTimer() { if (idx >= 0) cudaEventRecord(c->spans[idx].b, c->stream); }
};
}  // namespace rsvdb

#define RSVDB_CUDA(ctx, call)                                              \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return rsvdb::cuda_fail((ctx), e__, #call); } while (0)
#define RSVDB_TRY(call) do { int rc__ = (call); if (rc__ != 0) return rc__; } while (0)
