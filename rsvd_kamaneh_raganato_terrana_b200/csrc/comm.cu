#include "comm.cuh"
#include "context.cuh"

#include <dlfcn.h>
#include <cstring>

namespace rsvdb {

namespace {
// Minimal NCCL ABI (stable since 2.x): opaque comm, 128-byte unique id, enums as ints.
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_getuid)(nccl_uid*);
typedef int (*fn_initrank)(void** comm, int nranks, nccl_uid id, int rank);
typedef int (*fn_destroy)(void* comm);
typedef int (*fn_allreduce)(const void*, void*, size_t, int dtype, int op, void* comm, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int dtype, void* comm, cudaStream_t);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

struct Nccl {
  void* h = nullptr;
  fn_getuid getuid = nullptr; fn_initrank initrank = nullptr; fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr; fn_allgather allgather = nullptr; fn_errstr errstr = nullptr;
  bool ok = false; std::string why;
};
Nccl& nccl() {
  static Nccl n;
  if (n.h || !n.why.empty()) return n;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
  if (!n.h) { n.why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return n; }
  n.getuid = (fn_getuid)dlsym(n.h, "ncclGetUniqueId");
  n.initrank = (fn_initrank)dlsym(n.h, "ncclCommInitRank");
  n.destroy = (fn_destroy)dlsym(n.h, "ncclCommDestroy");
  n.allreduce = (fn_allreduce)dlsym(n.h, "ncclAllReduce");
  n.allgather = (fn_allgather)dlsym(n.h, "ncclAllGather");
  n.errstr = (fn_errstr)dlsym(n.h, "ncclGetErrorString");
  n.ok = n.getuid && n.initrank && n.destroy && n.allreduce && n.allgather;
  if (!n.ok) n.why = "libnccl is missing required symbols";
  return n;
}
int nccl_fail(rsvdb_ctx* c, int rc, const char* where) {
  Nccl& n = nccl();
  if (c) c->err = std::string(where) + ": " + (n.errstr ? n.errstr(rc) : "nccl error");
  return -3;
}
}  // namespace

int comm_unique_id(void* out128, std::string* err) {
  Nccl& n = nccl();
  if (!n.ok) { if (err) *err = n.why; return -3; }
  nccl_uid id; int rc = n.getuid(&id);
  if (rc != 0) { if (err) *err = n.errstr ? n.errstr(rc) : "ncclGetUniqueId failed"; return -3; }
  std::memcpy(out128, &id, 128);
  return 0;
}

int comm_init(rsvdb_ctx* c, int nranks, int rank, const void* id128) {
  Nccl& n = nccl();
  if (!n.ok) return fail(c, -3, n.why);
  if (c->nccl_comm) comm_destroy(c);
  nccl_uid id; std::memcpy(&id, id128, 128);
  void* comm = nullptr;
  int rc = n.initrank(&comm, nranks, id, rank);
  if (rc != 0) return nccl_fail(c, rc, "ncclCommInitRank");
  c->nccl_comm = comm; c->nranks = nranks; c->rank = rank;
  return 0;
}

void comm_destroy(rsvdb_ctx* c) {
  if (c && c->nccl_comm) { nccl().destroy(c->nccl_comm); c->nccl_comm = nullptr; c->nranks = 1; c->rank = 0; }
}

int comm_allreduce_sum(rsvdb_ctx* c, double* buf, size_t count) {
  if (c->nranks <= 1) return 0;
  int rc = nccl().allreduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, c->nccl_comm, c->stream);
  if (rc != 0) return nccl_fail(c, rc, "ncclAllReduce");
  return 0;
}

int comm_allreduce_max(rsvdb_ctx* c, double* buf, size_t count) {
  if (c->nranks <= 1) return 0;
  int rc = nccl().allreduce(buf, buf, count, NCCL_FLOAT64, NCCL_MAX, c->nccl_comm, c->stream);
  if (rc != 0) return nccl_fail(c, rc, "ncclAllReduce(max)");
  return 0;
}

int comm_allgather(rsvdb_ctx* c, const double* send, double* recv, size_t count) {
  if (c->nranks <= 1) {
    if (send != recv) { cudaError_t e = cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream); if (e != cudaSuccess) return cuda_fail(c, e, "allgather copy"); }
    return 0;
  }
  int rc = nccl().allgather(send, recv, count, NCCL_FLOAT64, c->nccl_comm, c->stream);
  if (rc != 0) return nccl_fail(c, rc, "ncclAllGather");
  return 0;
}

}  // namespace rsvdb
