// Uploads from PAGEABLE host memory at PCIe speed.
//
// The reference's callers hold their matrices in Eigen::MatrixXd -- ordinary pageable memory.  cudaMemcpyAsync from pageable
// memory goes through the driver's single bounce buffer (measured in round 1: ~11 GB/s, 70 of the 88 ms of a POD call on a
// 50000 x 2000 snapshot matrix).  HostStager owns a ring of pinned chunks and a few worker threads: the workers pack tile
// i+1 of the caller's matrix into a pinned chunk while the copy engine moves tile i, so the transfer runs at the rate of the
// slower of (parallel memcpy, PCIe) instead of the bounce-buffer rate.  Pinned sources bypass it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rsvdb {

class HostStager {
 public:
  HostStager();
  ~HostStager();
  // true when `p` is ordinary pageable memory (not cudaHostAlloc'ed / cudaHostRegister'ed / managed)
  static bool pageable(const void* p);
  // dst (device, column-major, leading dimension ldd) <- src (host, lds), rows x cols doubles, enqueued on `st`.
  // Returns when the last tile has been handed to the copy engine (the stream still has to be waited on, as for cudaMemcpyAsync).
  cudaError_t upload(cudaStream_t st, double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols);

 private:
  struct Impl;
  Impl* p_;
};

}  // namespace rsvdb
