// NCCL plumbing for row-sharded multi-GPU runs (one process per GPU).  libnccl.so.2 is resolved at run time with dlopen,
// so a single-GPU process never needs it and, under torch, the already-loaded torch-bundled NCCL is reused.
#pragma once
#include <cuda_runtime.h>
#include <string>

struct rsvdb_ctx;

namespace rsvdb {
int comm_unique_id(void* out128, std::string* err);
int comm_init(rsvdb_ctx* c, int nranks, int rank, const void* id128);
void comm_destroy(rsvdb_ctx* c);
// in-place sum over ranks of `count` doubles
int comm_allreduce_sum(rsvdb_ctx* c, double* buf, size_t count);
// in-place max over ranks of `count` doubles
int comm_allreduce_max(rsvdb_ctx* c, double* buf, size_t count);
// gather `count` doubles from every rank into recv (nranks * count), rank-major
int comm_allgather(rsvdb_ctx* c, const double* send, double* recv, size_t count);
}  // namespace rsvdb
