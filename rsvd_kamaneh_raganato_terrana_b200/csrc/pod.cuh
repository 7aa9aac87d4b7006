// POD wrappers around the SVD / rSVD path (reference POD/ParametricDiffusion1D/src/POD.cpp): correlation matrix, SVD
// dispatch, mode recovery and the energy criterion, with every product on the device.
#pragma once
#include <cstdint>
#include "context.cuh"

namespace rsvdb {

enum PodVariant { POD_NAIVE = 0, POD_STANDARD = 1, POD_ENERGY = 2, POD_WEIGHT = 3 };

struct PodShape { int64_t w_cols_full; int64_t sigma_len; };
// Shapes the reference would produce BEFORE the energy truncation (W columns) and for sigma, per variant / svd_type
// (POD.cpp:42-114,116-134,136-224).  Returns false for an svd_type outside [0,5].
bool pod_shape(int variant, int64_t Nh, int64_t ns, int r, int svd_type, PodShape* out);

// S (Nh x ns, device).  Xh (Nh x Nh) for POD_ENERGY / POD_WEIGHT, D (ns x ns) for POD_WEIGHT.  Omega (optional, device):
// sketch for svd_type 3-5, (columns of the matrix handed to rSVD) x r.  Outputs on the device: W (Nh x w_cols_full, ldw),
// sigma (sigma_len); *N = number of modes the energy criterion keeps (POD.cpp:203-219); for POD_NAIVE *N = w_cols_full.
// Synchronises the stream once (the criterion is evaluated on the host, like the reference).
int pod_device(rsvdb_ctx* c, int variant, const double* S, int64_t Nh, int64_t ns, int64_t lds, const double* Xh, int64_t ldx,
               const double* D, int64_t ldd, int r, double tol, int svd_type, uint64_t seed, const double* Omega, int64_t ldo,
               double* W, int64_t ldw, double* sigma, int* N);

}  // namespace rsvdb
