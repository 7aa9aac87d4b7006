// Per-device "already done" flag for cudaFuncSetAttribute: function attributes are per device, a process may hold
// contexts on several GPUs (rsvdb_create(&c, device)), so a process-wide static bool is not enough.
#pragma once
#include <cuda_runtime.h>

namespace rsvdb {
struct DevOnce {
  bool done[64] = {};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool get() const { return done[dev()]; }
  void set() { done[dev()] = true; }
};
}  // namespace rsvdb
