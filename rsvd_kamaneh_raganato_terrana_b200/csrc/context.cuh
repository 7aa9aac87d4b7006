// Engine context behind the C ABI (include/rsvdb.h).
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>
#include "gemm_dmma.cuh"

struct rsvdb_ctx {
  int device = 0;
  int nsm = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  rsvdb::GemmWorkspace gemm_ws;     // split-K partial tiles
  rsvdb::GemmWorkspace qr_ws;       // TSQR tree storage
  rsvdb::GemmWorkspace tmp_ws;      // pipeline intermediates (Y, Z, B, ...)
  rsvdb::GemmWorkspace io_ws;       // device copies for the *_host entry points
  int launches_i = 0;               // bumped by the launchers
  int64_t launches = 0;
  std::string err;
  // multi-GPU (row-sharded A); comm is an ncclComm_t resolved at run time (comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
};

namespace rsvdb {
inline int fail(rsvdb_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
inline int cuda_fail(rsvdb_ctx* c, cudaError_t e, const char* where) {
  if (c) c->err = std::string(where) + ": " + cudaGetErrorString(e);
  return -2;
}
struct DeviceGuard {
  int prev = -1; bool ok;
  explicit DeviceGuard(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace rsvdb

#define RSVDB_CUDA(ctx, call)                                              \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return rsvdb::cuda_fail((ctx), e__, #call); } while (0)
