// C ABI (include/rsvdb.h): context management and the dense building blocks.
#include "../../include/rsvdb.h"
#include "context.cuh"

using namespace rsvdb;

extern "C" {

const char* rsvdb_version(void) { return "rsvdb 0.1 (sm_100a)"; }

int rsvdb_create(rsvdb_ctx** out, int device) {
  if (!out) return RSVDB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return RSVDB_ERR_CUDA;
  rsvdb_ctx* c = new rsvdb_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  if (prop.major != 10) { delete c; return RSVDB_ERR_UNSUPPORTED; }   // sm_100a only: no other code path exists
  c->nsm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RSVDB_ERR_CUDA; }
  c->stream = c->own_stream;
  *out = c;
  return RSVDB_OK;
}

int rsvdb_destroy(rsvdb_ctx* c) {
  if (!c) return RSVDB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->gemm_ws.release(); c->qr_ws.release(); c->tmp_ws.release(); c->io_ws.release();
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return RSVDB_OK;
}

int rsvdb_set_stream(rsvdb_ctx* c, void* s) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  c->stream = static_cast<cudaStream_t>(s);
  return RSVDB_OK;
}

int rsvdb_use_own_stream(rsvdb_ctx* c) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  c->stream = c->own_stream;
  return RSVDB_OK;
}

int rsvdb_synchronize(rsvdb_ctx* c) {
  if (!c) return RSVDB_ERR_INVALID_ARGUMENT;
  RSVDB_CUDA(c, cudaStreamSynchronize(c->stream));
  return RSVDB_OK;
}

const char* rsvdb_last_error(const rsvdb_ctx* c) { return c ? c->err.c_str() : "null context"; }
int64_t rsvdb_launch_count(const rsvdb_ctx* c) { return c ? c->launches : 0; }

int rsvdb_gemm_an_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dX, int64_t ldx,
                      int l, double* dY, int64_t ldy) {
  if (!c || m < 0 || n < 0 || l < 0 || lda < m || ldx < n || ldy < m) return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "gemm_an: bad shape");
  int k = 0;
  RSVDB_CUDA(c, gemm_an(c->gemm_ws, c->stream, c->nsm, dA, m, n, lda, dX, ldx, l, dY, ldy, &k));
  c->launches += k;
  return RSVDB_OK;
}

int rsvdb_gemm_at_dev(rsvdb_ctx* c, const double* dA, int64_t m, int64_t n, int64_t lda, const double* dQ, int64_t ldq,
                      int l, double* dZ, int64_t ldz, int transpose_out) {
  if (!c || m < 0 || n < 0 || l < 0 || lda < m || ldq < m || ldz < (transpose_out ? l : n))
    return fail(c, RSVDB_ERR_INVALID_ARGUMENT, "gemm_at: bad shape");
  int k = 0;
  RSVDB_CUDA(c, gemm_at(c->gemm_ws, c->stream, c->nsm, dA, m, n, lda, dQ, ldq, l, dZ, ldz, transpose_out, &k));
  c->launches += k;
  return RSVDB_OK;
}

}  // extern "C"
