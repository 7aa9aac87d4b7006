/* TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load this.  Nothing under rsvd_kamaneh_raganato_terrana_b200/ may.
 *
 * Plain-C restatement of the loop-heavy first-party pieces of the reference's rSVD path.  All matrices are
 * column-major FP64, `ld` = rows.  Each function cites the reference file:line it follows.  The restatement is
 * checked against the reference's own sources (compiled over oracle/eigen_shim into oracle/_ref) by
 * tests/test_oracle_vs_ref.py in the dev container, and against tests/golden/ everywhere.
 *
 * Pin status: the reference holds no golden vectors for this path, so the pin is
 * "outputs of the reference's first-party code run here over an Eigen/MPI stand-in" plus the mathematical
 * known answers of BASELINE.md section 3.  The Eigen arithmetic itself (GEMM, HouseholderQR) is restated, not run.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AT(M, ld, i, j) ((M)[(size_t)(i) + (size_t)(j) * (size_t)(ld)])

/* ---- Householder QR: Eigen::HouseholderQR semantics (third-party; restated from Golub & Van Loan Alg. 5.2.1 with
 * the sign convention beta = -sign(x0)*||x||, sign(0)=+1; call sites src/rSVD.cpp:60-68, include/SVD_class.hpp:112-122).
 * On return F holds R in its upper triangle and the essential reflector parts below; tau has min(m,n) entries. */
void oc_householder_qr(double* F, long m, long n, double* tau) {
  long k = m < n ? m : n;
  for (long j = 0; j < k; ++j) {
    double tail = 0.0;
    for (long i = j + 1; i < m; ++i) tail += AT(F, m, i, j) * AT(F, m, i, j);
    double x0 = AT(F, m, j, j), beta, t;
    if (tail <= DBL_MIN) {
      t = 0.0; beta = x0;
      for (long i = j + 1; i < m; ++i) AT(F, m, i, j) = 0.0;
    } else {
      beta = sqrt(x0 * x0 + tail);
      if (x0 >= 0.0) beta = -beta;
      double inv = 1.0 / (x0 - beta);
      for (long i = j + 1; i < m; ++i) AT(F, m, i, j) *= inv;
      t = (beta - x0) / beta;
    }
    AT(F, m, j, j) = beta; tau[j] = t;
    if (t != 0.0)
      for (long c = j + 1; c < n; ++c) {
        double w = AT(F, m, j, c);
        for (long i = j + 1; i < m; ++i) w += AT(F, m, i, j) * AT(F, m, i, c);
        w *= t;
        AT(F, m, j, c) -= w;
        for (long i = j + 1; i < m; ++i) AT(F, m, i, c) -= w * AT(F, m, i, j);
      }
  }
}
/* C (m x nc) <- H_0 H_1 ... H_{k-1} C  (householderQ() * C). */
void oc_householder_apply_q(const double* F, long m, long k, const double* tau, double* C, long nc) {
  for (long j = k - 1; j >= 0; --j) {
    if (tau[j] == 0.0) continue;
    for (long c = 0; c < nc; ++c) {
      double w = AT(C, m, j, c);
      for (long i = j + 1; i < m; ++i) w += AT(F, m, i, j) * AT(C, m, i, c);
      w *= tau[j];
      AT(C, m, j, c) -= w;
      for (long i = j + 1; i < m; ++i) AT(C, m, i, c) -= w * AT(F, m, i, j);
    }
  }
}

/* ---- plane rotations: src/JacobiOperations.cpp:6-24 ---- */
static void rot_left(double* M, long ld, long ncols, long p, long q, double c, double s) {   /* :6-14 */
  for (long i = 0; i < ncols; ++i) {
    double x = AT(M, ld, p, i), y = AT(M, ld, q, i);
    AT(M, ld, p, i) = c * x + s * y;
    AT(M, ld, q, i) = -s * x + c * y;
  }
}
static void rot_right(double* M, long ld, long nrows, long p, long q, double c, double s) {  /* :16-24 */
  for (long i = 0; i < nrows; ++i) {
    double x = AT(M, ld, i, p), y = AT(M, ld, i, q);
    AT(M, ld, i, p) = c * x + (-s) * y;
    AT(M, ld, i, q) = s * x + c * y;
  }
}

/* JacobiRotation::makeJacobi: src/Jacobi_Class.cpp:39-60.  Returns 1 if a rotation was produced. */
int oc_make_jacobi(double x, double y, double z, double* c, double* s) {
  double deno = 2.0 * fabs(y);
  if (deno < DBL_MIN) { *c = 1.0; *s = 0.0; return 0; }
  double tau = (x - z) / deno, w = sqrt(tau * tau + 1.0);
  double t = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sgn = t > 0 ? 1.0 : -1.0, nn = 1.0 / sqrt(t * t + 1.0);
  *s = -sgn * (y / fabs(y)) * fabs(t) * nn; *c = nn;
  return 1;
}

/* real_2x2_jacobi_svd: src/JacobiOperations.cpp:25-88 (deno_floor = DBL_MIN) and its _par twin :140-203
 * (deno_floor = 1e-10, :168).  m = [m00 m01; m10 m11]. */
void oc_real_2x2_jacobi_svd(double m00, double m01, double m10, double m11, double deno_floor,
                            double* cl, double* sl, double* cr, double* sr) {
  double t = m00 + m11, d = m10 - m01, rc, rs;
  if (d == 0.0) { rs = 0.0; rc = 1.0; }
  else { double u = t / d, tmp = sqrt(1.0 + u * u); rs = 1.0 / tmp; rc = u / tmp; }
  /* rows (0,1) <- rot1 applied on the left (:45) */
  double a00 = rc * m00 + rs * m10, a01 = rc * m01 + rs * m11;
  double a11 = -rs * m01 + rc * m11;
  double deno = 2.0 * fabs(a01);
  if (deno < deno_floor) { *cr = 1.0; *sr = 0.0; }
  else {
    double tau = (a00 - a11) / deno, w = sqrt(tau * tau + 1.0);
    double t2 = tau > 0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
    double sgn = t2 > 0 ? 1.0 : -1.0, nn = 1.0 / sqrt(t2 * t2 + 1.0);
    *sr = -sgn * (a01 / fabs(a01)) * fabs(t2) * nn; *cr = nn;
  }
  /* left = rot1 * j_right^T with j_right = [cr sr; -sr cr] (:73-84) */
  *cl = rc * (*cr) + rs * (*sr);
  *sl = rc * (-(*sr)) + rs * (*cr);
}

/* QR preconditioner shared by jacobiSVD (include/SVD_class.hpp:110-123) and ParallelJacobiSVD (:238-250).
 * Produces work (k x k), U (m x k), V (n x k), k = min(m,n). */
static void precondition(const double* A, long m, long n, double* W, double* U, double* V) {
  long k = m < n ? m : n;
  memset(U, 0, sizeof(double) * (size_t)m * (size_t)k);
  memset(V, 0, sizeof(double) * (size_t)n * (size_t)k);
  for (long i = 0; i < k; ++i) { AT(U, m, i, i) = 1.0; AT(V, n, i, i) = 1.0; }
  if (m > n) {
    double* F = (double*)malloc(sizeof(double) * (size_t)m * (size_t)n); double* tau = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(F, A, sizeof(double) * (size_t)m * (size_t)n);
    oc_householder_qr(F, m, n, tau);
    for (long j = 0; j < n; ++j) for (long i = 0; i < n; ++i) AT(W, n, i, j) = (i <= j) ? AT(F, m, i, j) : 0.0;
    oc_householder_apply_q(F, m, n, tau, U, n);
    free(F); free(tau);
  } else if (n > m) {
    double* F = (double*)malloc(sizeof(double) * (size_t)m * (size_t)n); double* tau = (double*)malloc(sizeof(double) * (size_t)m);
    for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) AT(F, n, j, i) = AT(A, m, i, j);   /* adjoint, n x m */
    oc_householder_qr(F, n, m, tau);
    for (long j = 0; j < m; ++j) for (long i = 0; i < m; ++i) AT(W, m, j, i) = (i <= j) ? AT(F, n, i, j) : 0.0;  /* R^T */
    oc_householder_apply_q(F, n, m, tau, V, m);
    free(F); free(tau);
  } else {
    memcpy(W, A, sizeof(double) * (size_t)m * (size_t)n);
  }
}

/* abs / sign-fix / selection sort: include/SVD_class.hpp:158-178 (= :310-332). */
static void finish(double* W, long k, double* U, long m, double* V, long n, double* S) {
  for (long i = 0; i < k; ++i) {
    double a = AT(W, k, i, i); S[i] = fabs(a);
    if (a < 0) for (long r = 0; r < m; ++r) AT(U, m, r, i) = -AT(U, m, r, i);
  }
  for (long i = 0; i < k; ++i) {
    long pos = 0; double best = S[i];
    for (long j = 1; j < k - i; ++j) if (S[i + j] > best) { best = S[i + j]; pos = j; }
    if (best == 0) break;
    if (pos) {
      pos += i;
      double t = S[i]; S[i] = S[pos]; S[pos] = t;
      for (long r = 0; r < m; ++r) { double x = AT(U, m, r, pos); AT(U, m, r, pos) = AT(U, m, r, i); AT(U, m, r, i) = x; }
      for (long r = 0; r < n; ++r) { double x = AT(V, n, r, pos); AT(V, n, r, pos) = AT(V, n, r, i); AT(V, n, r, i) = x; }
    }
  }
}

static int precond_is_real(const double* W, long k, long p, long q, double maxDiag) {   /* JacobiOperations.cpp:89-103 */
  return !(fabs(AT(W, k, p, q)) < maxDiag * DBL_EPSILON && fabs(AT(W, k, q, p)) < maxDiag * DBL_EPSILON);
}

/* SVD<Jacobi>::jacobiSVD: include/SVD_class.hpp:101-180.  U m x k, S k, V n x k.  Returns the number of sweeps;
 * *rotations (optional) counts the 2x2 SVDs applied. */
long oc_jacobi_svd(const double* A, long m, long n, double* U, double* S, double* V, long* rotations) {
  long k = m < n ? m : n, sweeps = 0, rots = 0;
  double* W = (double*)calloc((size_t)k * (size_t)k, sizeof(double));
  precondition(A, m, n, W, U, V);
  const double considerAsZero = DBL_MIN, precision = 2.0 * DBL_EPSILON;
  double maxDiag = 0.0;
  for (long i = 0; i < k; ++i) if (fabs(AT(W, k, i, i)) > maxDiag) maxDiag = fabs(AT(W, k, i, i));
  int finished = 0;
  while (!finished) {
    finished = 1; ++sweeps;
    for (long p = 1; p < k; ++p)
      for (long q = 0; q < p; ++q) {
        double thr = fmax(considerAsZero, precision * maxDiag);
        if (fabs(AT(W, k, p, q)) > thr || fabs(AT(W, k, q, p)) > thr) {
          finished = 0;
          if (precond_is_real(W, k, p, q, maxDiag)) {
            double cl, sl, cr, sr;
            oc_real_2x2_jacobi_svd(AT(W, k, p, p), AT(W, k, p, q), AT(W, k, q, p), AT(W, k, q, q), DBL_MIN, &cl, &sl, &cr, &sr);
            rot_left(W, k, k, p, q, cl, sl);
            rot_right(U, m, m, p, q, cl, -sl);
            rot_right(W, k, k, p, q, cr, sr);
            rot_right(V, n, n, p, q, cr, sr);
            maxDiag = fmax(maxDiag, fmax(fabs(AT(W, k, p, p)), fabs(AT(W, k, q, q))));
            ++rots;
          }
        }
      }
  }
  finish(W, k, U, m, V, n, S);
  free(W);
  if (rotations) *rotations = rots;
  return sweeps;
}

typedef struct { double w; long p, q; } wpq;
static int wpq_desc(const void* a, const void* b) {   /* std::sort(..., std::greater<>()) on (weight,p,q) tuples: SVD_class.hpp:286 */
  const wpq* x = (const wpq*)a; const wpq* y = (const wpq*)b;
  if (x->w != y->w) return x->w > y->w ? -1 : 1;
  if (x->p != y->p) return x->p > y->p ? -1 : 1;
  if (x->q != y->q) return x->q > y->q ? -1 : 1;
  return 0;
}
/* SVD<ParallelJacobi>::ParallelJacobiSVD: include/SVD_class.hpp:224-333 (OpenMP there only splits the pair scan and the
 * per-rotation loops; the arithmetic order per element is unchanged). */
long oc_parallel_jacobi_svd(const double* A, long m, long n, double* U, double* S, double* V, long* rotations) {
  long k = m < n ? m : n, passes = 0, rots = 0;
  double* W = (double*)calloc((size_t)k * (size_t)k, sizeof(double));
  precondition(A, m, n, W, U, V);
  const double considerAsZero = 1e-12, precision = 1e-12;
  double maxDiag = 0.0;
  for (long i = 0; i < k; ++i) if (fabs(AT(W, k, i, i)) > maxDiag) maxDiag = fabs(AT(W, k, i, i));
  wpq* list = (wpq*)malloc(sizeof(wpq) * (size_t)(k * (k - 1) / 2 + 1));
  int finished = 0;
  while (!finished) {
    finished = 1; ++passes;
    long cnt = 0;
    for (long p = 1; p < k; ++p)
      for (long q = 0; q < p; ++q) {
        double thr = fmax(considerAsZero, precision * maxDiag);
        double w = AT(W, k, p, q) * AT(W, k, p, q) + AT(W, k, q, p) * AT(W, k, q, p);
        if (w > thr) { finished = 0; list[cnt].w = w; list[cnt].p = p; list[cnt].q = q; ++cnt; }
      }
    if (cnt) {
      qsort(list, (size_t)cnt, sizeof(wpq), wpq_desc);
      for (long e = 0; e < cnt; ++e) {
        long p = list[e].p, q = list[e].q;
        if (precond_is_real(W, k, p, q, maxDiag)) {
          double cl, sl, cr, sr;
          oc_real_2x2_jacobi_svd(AT(W, k, p, p), AT(W, k, p, q), AT(W, k, q, p), AT(W, k, q, q), 1e-10, &cl, &sl, &cr, &sr);
          rot_left(W, k, k, p, q, cl, sl);
          rot_right(U, m, m, p, q, cl, -sl);
          rot_right(W, k, k, p, q, cr, sr);
          rot_right(V, n, n, p, q, cr, sr);
          maxDiag = fmax(maxDiag, fmax(fabs(AT(W, k, p, p)), fabs(AT(W, k, q, q))));
          ++rots;
        }
      }
    }
  }
  finish(W, k, U, m, V, n, S);
  free(W); free(list);
  if (rotations) *rotations = rots;
  return passes;
}

/* PM iteration count: src/PM.cpp:25-28. */
int oc_pm_iterations(long ncols) {
  double epsilon = 1.e-10, delta = 0.05, lambda = 0.1;
  return (int)ceil(log(4 * log(2 * ncols / delta) / (epsilon * delta)) / (2 * lambda));
}
/* PM(A, B, sigma, u, v): src/PM.cpp:4-81 at one MPI rank, with the start vector x0 (length n, any non-zero) supplied by
 * the caller -- the reference draws it from std::random_device (:15-21), so it is not a parity surface. */
void oc_pm(const double* A, long m, long n, const double* B, double* x0, double* sigma, double* u, double* v) {
  double* res = (double*)malloc(sizeof(double) * (size_t)n);
  double nr = 0; for (long i = 0; i < n; ++i) nr += x0[i] * x0[i]; nr = sqrt(nr);
  for (long i = 0; i < n; ++i) x0[i] /= nr;
  int s = oc_pm_iterations(n);
  for (int it = 1; it <= s; ++it) {
    for (long i = 0; i < n; ++i) { double acc = 0.0; for (long j = 0; j < n; ++j) acc += AT(B, n, i, j) * x0[j]; res[i] = acc; }
    nr = 0; for (long i = 0; i < n; ++i) nr += res[i] * res[i]; nr = sqrt(nr);
    for (long i = 0; i < n; ++i) x0[i] = nr > 0 ? res[i] / nr : res[i];
  }
  nr = 0; for (long i = 0; i < n; ++i) nr += x0[i] * x0[i]; nr = sqrt(nr);
  for (long i = 0; i < n; ++i) v[i] = nr > 0 ? x0[i] / nr : x0[i];
  double sg = 0;
  for (long i = 0; i < m; ++i) { double acc = 0; for (long j = 0; j < n; ++j) acc += AT(A, m, i, j) * v[j]; u[i] = acc; sg += acc * acc; }
  sg = sqrt(sg); *sigma = sg;
  for (long i = 0; i < m; ++i) u[i] /= sg;
  free(res);
}
/* SVD<Power>::powerMethodSVD: include/SVD_class.hpp:184-219.  U is m x m (identity-initialised by compute(), :82), V is
 * n x n with right singular vectors in ROWS (:214), S has min(m,n) entries; returns the number of triplets found
 * (the reference conservativeResize()s on early exit, :198-209).  starts: dim start vectors of length n, column-major. */
long oc_power_svd(const double* A_in, long m, long n, int r, const double* starts, double* U, double* S, double* V) {
  long k = m < n ? m : n, dim = r ? r : k;
  double* A = (double*)malloc(sizeof(double) * (size_t)m * (size_t)n); memcpy(A, A_in, sizeof(double) * (size_t)m * (size_t)n);
  double* B = (double*)calloc((size_t)n * (size_t)n, sizeof(double));
  for (long j = 0; j < n; ++j) for (long i = 0; i < n; ++i) { double acc = 0; for (long t = 0; t < m; ++t) acc += AT(A, m, t, i) * AT(A, m, t, j); AT(B, n, i, j) = acc; }
  memset(U, 0, sizeof(double) * (size_t)m * (size_t)m); for (long i = 0; i < m; ++i) AT(U, m, i, i) = 1.0;
  memset(V, 0, sizeof(double) * (size_t)n * (size_t)n); for (long i = 0; i < n; ++i) AT(V, n, i, i) = 1.0;
  memset(S, 0, sizeof(double) * (size_t)k);
  double* u = (double*)malloc(sizeof(double) * (size_t)m); double* v = (double*)malloc(sizeof(double) * (size_t)n);
  double* x0 = (double*)malloc(sizeof(double) * (size_t)n);
  long found = 0;
  for (long i = 0; i < dim; ++i) {
    double sigma; memcpy(x0, starts + (size_t)i * (size_t)n, sizeof(double) * (size_t)n);
    oc_pm(A, m, n, B, x0, &sigma, u, v);
    if (sigma < 1e-12) break;
    /* update = sigma u v^T; data -= update; B -= update^T update = sigma^2 v v^T (u has unit norm up to rounding; the
       reference forms update^T*update explicitly, :210-212, and so does this) */
    double uu = 0; for (long t = 0; t < m; ++t) uu += (sigma * u[t]) * (sigma * u[t]);
    for (long c = 0; c < n; ++c) for (long t = 0; t < m; ++t) AT(A, m, t, c) -= sigma * u[t] * v[c];
    for (long c = 0; c < n; ++c) for (long t = 0; t < n; ++t) AT(B, n, t, c) -= v[t] * uu * v[c];
    for (long t = 0; t < m; ++t) AT(U, m, t, i) = u[t];
    for (long t = 0; t < n; ++t) AT(V, n, i, t) = v[t];
    S[i] = sigma; ++found;
  }
  free(A); free(B); free(u); free(v); free(x0);
  return found;
}

/* Givens QR: src/QR.cpp:12-80 (= image_compression/src/QR.cpp:45-99).  full: Q m x m, R m x n.  reduced additionally
 * truncates to Q[:, :n], R[:n, :] (:78-79) -- done by the caller. */
void oc_givens_qr_full(const double* A, long m, long n, double* Q, double* R) {
  memset(Q, 0, sizeof(double) * (size_t)m * (size_t)m); for (long i = 0; i < m; ++i) AT(Q, m, i, i) = 1.0;
  memcpy(R, A, sizeof(double) * (size_t)m * (size_t)n);
  long k = m < n ? m : n;
  for (long j = 0; j < k; ++j)
    for (long i = m - 1; i > j; --i)
      if (AT(R, m, i, j) != 0) {
        double a = AT(R, m, i - 1, j), b = AT(R, m, i, j), rr = hypot(a, b), c = a / rr, s = -b / rr;   /* G = [c -s; s c] */
        for (long t = j; t < n; ++t) {
          double x = AT(R, m, i - 1, t), y = AT(R, m, i, t);
          AT(R, m, i - 1, t) = c * x + (-s) * y;
          AT(R, m, i, t) = s * x + c * y;
        }
        for (long t = 0; t < m; ++t) {   /* Q[:, i-1:i+1] *= G^T */
          double x = AT(Q, m, t, i - 1), y = AT(Q, m, t, i);
          AT(Q, m, t, i - 1) = x * c + y * (-s);
          AT(Q, m, t, i) = x * s + y * c;
        }
      }
}

/* manualMatrixMultiply: src/matrixOperations.cpp:7-28.  Returns -1 on a shape mismatch (the reference throws). */
int oc_manual_matmul(const double* A, long m, long ka, const double* B, long kb, long n, double* C) {
  if (ka != kb) return -1;
  for (long i = 0; i < m; ++i) for (long j = 0; j < n; ++j) { double s = 0; for (long t = 0; t < ka; ++t) s += AT(A, m, i, t) * AT(B, kb, t, j); AT(C, m, i, j) = s; }
  return 0;
}
