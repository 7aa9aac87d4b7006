// TSQR driver (see tsqr.cu).
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include "gemm_dmma.cuh"

namespace rsvdb {

class Tsqr {
 public:
  explicit Tsqr(GemmWorkspace* ws, cudaStream_t side = nullptr, cudaEvent_t* events = nullptr) : ws_(ws), side_(side), ev_(events) {}
  // Size the tree for a rows x l panel and reserve workspace.
  cudaError_t plan(long long rows, int l);
  // Factor Y in place (reflectors overwrite Y).  Afterwards R_local() is the l x l upper-triangular factor of this panel.
  cudaError_t factor(cudaStream_t st, double* Y, long long ldy, int* launches);
  // Overwrite Y with the explicit thin Q.  Ctop (l x l, ldc) multiplies from the right at the top of the tree:
  // Q = Q_tree * Ctop; nullptr = identity.
  cudaError_t form_q(cudaStream_t st, double* Y, long long ldy, const double* Ctop, long long ldc, int* launches);
  const double* R_local() const;
  double* top_block();                 // l x l scratch owned by the plan (for the multi-GPU top block)
  int leaf_rows() const { return br_; }
  int depth() const { return (int)levels_.size(); }

 private:
  struct Level { long long rows; int nb; size_t off_R; size_t off_tau; size_t off_T; size_t off_E; int cl; int node_rows; };   // cl: CTAs per cluster node (0 = single-CTA blocks)
  GemmWorkspace* ws_;
  cudaStream_t side_ = nullptr;        // optional: upper-level explicit factors are formed here, overlapping the factor chain
  cudaEvent_t* ev_ = nullptr;          // >= 17 events
  bool upper_done_ = false;
  std::vector<Level> levels_;
  long long rows_ = 0;
  int l_ = 0, br_ = 0;
  bool blk_ = false;                   // blocked (compact-WY, DMMA) leaves
  size_t off_top_ = 0, off_scratch_ = 0;
};

// Householder full QR through global memory (small inputs): F factored in place, Rsq = cols x cols upper triangle,
// Q = rows x rows.  tau: cols doubles.
cudaError_t house_full_qr(cudaStream_t st, double* F, long long ldf, long long rows, int cols, double* tau, double* Rsq, double* Q,
                          long long ldq, int* launches);

}  // namespace rsvdb
