"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's rSVD hot path.  NOT product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs (``cpu_baseline`` / ``--impl reference``)
may import this module.  The product package ``rsvd_kamaneh_raganato_terrana_b200`` never does.

What is restated (reference file:line, relative to /root/reference):

* ``intermediate_step``  src/rSVD.cpp:57-70   (range finder with q power iterations, QR after every product)
* ``rsvd``               src/rSVD.cpp:72-133  (B = Q^T A, small SVD dispatch, U = Q*Utilde); Omega is an ARGUMENT
                                               here (the reference draws it from std::random_device, :26-28)
* small SVD back-ends    include/SVD_class.hpp:101-180 (Jacobi), :224-333 (ParallelJacobi), :184-219 + src/PM.cpp
                         (Power) -- the loops live in oracle/oracle_c.c
* Givens QR              src/QR.cpp:12-80, manualMatrixMultiply src/matrixOperations.cpp:7-28 -- oracle_c.c
* ``pod``                POD/ParametricDiffusion1D/src/POD.cpp:42-114 (perform_SVD), :116-134 (naive), :136-224
                         (standard), :226-336 (energy), :338-461 (weight); Eigen's SelfAdjointEigenSolver::operatorSqrt
                         and ConjugateGradient are restated through numpy eigh / solve
* ``image_*``            image_compression/src/image_com.cpp:184-190,251-317 (normalise, compress = older-API rSVD with
                         l = k + 10, reconstruct, deNormalize); NOT compiled from the reference (its Image class is only
                         reachable through the stb PNG codec): "parity unpinned" for this wrapper, the pieces are pinned above
* ``PCA``                PCA/include/PCA_class.hpp:24-47 (centre, optional stddev scaling, SVD<method>), :76-100
                         (explained variance / ratio, scores, loadings, projectToPCA, reconstructFromPCA)

Third-party arithmetic: the reference calls Eigen (un-vendored, un-pinned: Makefile:2 ``-I ${mkEigenInc}``) for the dense
products and ``Eigen::HouseholderQR``.  Eigen is absent from this image.  Its published algorithm (unblocked/blocked
Householder QR with beta = -sign(x0)*||x||, identical to LAPACK dlarfg/dgeqrf/dorgqr) is restated through
``scipy.linalg.qr(mode="economic")`` for large inputs and ``oc_householder_qr`` (oracle_c.c) for small ones; both give the
same Q, R up to rounding.  Products go through numpy (OpenBLAS dgemm).

Pin status: the reference holds NO golden vectors for this path.  This oracle is pinned instead against
(1) outputs of the reference's own first-party sources compiled here over an Eigen/MPI stand-in (oracle/_ref, built by
oracle/Makefile; compared in tests/test_oracle_vs_ref.py and frozen as tests/golden/*.npz by
tests/golden/make_golden.py), and (2) the mathematical known answers of BASELINE.md section 3.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import scipy.linalg

_HERE = Path(__file__).resolve().parent
_LIB = None

JACOBI, POWER, PARALLEL_JACOBI = 0, 1, 2   # enum class SVDMethod, include/SVD_class.hpp:28-32


def build(force: bool = False) -> Path:
    """Compile oracle_c.c (and, when /root/reference is present, oracle/_ref) with the committed Makefile."""
    so = _HERE / "_build" / "liboracle_c.so"
    src = _HERE / "oracle_c.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "oracle_c"], check=True, capture_output=True)
    ref_so = _HERE / "_ref" / "libref_rsvd.so"
    drv = _HERE / "ref_driver.cpp"
    if Path("/root/reference/src").is_dir() and (force or not ref_so.exists() or ref_so.stat().st_mtime < drv.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "ref"], check=True, capture_output=True)
    ref1_so = _HERE / "_ref" / "libref_imgcomp.so"
    drv1 = _HERE / "ref_driver_v1.cpp"
    if Path("/root/reference/image_compression/src").is_dir() and (force or not ref1_so.exists() or ref1_so.stat().st_mtime < drv1.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "ref_v1"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.oc_jacobi_svd.restype = ctypes.c_long
        _LIB.oc_parallel_jacobi_svd.restype = ctypes.c_long
        _LIB.oc_power_svd.restype = ctypes.c_long
    return _LIB


_dp = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------------
# Householder QR (Eigen::HouseholderQR semantics)
# ---------------------------------------------------------------------------------------------------------------
def householder_qr(Y, small_limit: int = 200_000):
    """Return (Q_thin, R) with Q_thin = householderQ() * Identity(rows, cols)  (src/rSVD.cpp:60-61)."""
    Y = _f(Y)
    m, n = Y.shape
    if m * n <= small_limit:
        F = Y.copy(order="F")
        k = min(m, n)
        tau = np.zeros(k)
        _lib().oc_householder_qr(_p(F), ctypes.c_long(m), ctypes.c_long(n), _p(tau))
        Q = np.asfortranarray(np.eye(m, n))
        _lib().oc_householder_apply_q(_p(F), ctypes.c_long(m), ctypes.c_long(k), _p(tau), _p(Q), ctypes.c_long(n))
        return Q, np.triu(F[:k, :])
    Q, R = scipy.linalg.qr(Y, mode="economic", overwrite_a=False, check_finite=False)
    return _f(Q), R


def intermediate_step(A, Omega, l: int, q: int):
    """src/rSVD.cpp:57-70.  Returns Q (m x l)."""
    A = np.asarray(A, dtype=np.float64)
    Y = A @ Omega                               # :59
    Q, _ = householder_qr(Y)                    # :60-61
    for _ in range(q):                          # :62
        Y = A.T @ Q                             # :63
        Q, _ = householder_qr(Y)                # :64-65
        Y = A @ Q                               # :66
        Q, _ = householder_qr(Y)                # :67-68
    return Q[:, :l]


# ---------------------------------------------------------------------------------------------------------------
# Small SVD back-ends (SVD<method>::compute)
# ---------------------------------------------------------------------------------------------------------------
def svd_jacobi(B):
    """include/SVD_class.hpp:101-180.  Returns (U m x k, S k, V n x k, info)."""
    B = _f(B); m, n = B.shape; k = min(m, n)
    U = np.zeros((m, k), order="F"); S = np.zeros(k); V = np.zeros((n, k), order="F"); rot = ctypes.c_long(0)
    sweeps = _lib().oc_jacobi_svd(_p(B), ctypes.c_long(m), ctypes.c_long(n), _p(U), _p(S), _p(V), ctypes.byref(rot))
    return U, S, V, {"sweeps": int(sweeps), "rotations": int(rot.value)}


def svd_parallel_jacobi(B):
    """include/SVD_class.hpp:224-333."""
    B = _f(B); m, n = B.shape; k = min(m, n)
    U = np.zeros((m, k), order="F"); S = np.zeros(k); V = np.zeros((n, k), order="F"); rot = ctypes.c_long(0)
    passes = _lib().oc_parallel_jacobi_svd(_p(B), ctypes.c_long(m), ctypes.c_long(n), _p(U), _p(S), _p(V), ctypes.byref(rot))
    return U, S, V, {"passes": int(passes), "rotations": int(rot.value)}


def svd_power(B, r: int = 0, seed: int = 0, resize: bool = True):
    """include/SVD_class.hpp:184-219 + src/PM.cpp:4-81.  Start vectors come from a seeded generator (the reference uses
    std::random_device).  Returns (U m x m [or m x found on early exit], S, V n x n with singular vectors in ROWS).
    resize=False: the OLDER API's singularValueDecomposition (image_compression/src/SVD.cpp:30-55) has no sigma < 1e-12 early
    exit and never resizes its outputs; for it the triplets past the numerical rank are left at their initial values
    (identity columns / rows, sigma 0) -- the reference computes rounding noise there (sigma ~ 1e-16)."""
    B = _f(B); m, n = B.shape; k = min(m, n); dim = r if r else k
    starts = _f(np.random.default_rng(seed).standard_normal((n, dim)))
    U = np.zeros((m, m), order="F"); S = np.zeros(k); V = np.zeros((n, n), order="F")
    found = int(_lib().oc_power_svd(_p(B), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(r), _p(starts), _p(U), _p(S), _p(V)))
    if found < dim and resize:    # conservativeResize on early exit, :198-209
        if found == 0:
            return np.zeros((m, 1), order="F"), np.zeros(1), np.zeros((n, 1), order="F"), {"found": 0}
        return _f(U[:, :found]), S[:found].copy(), _f(V[:, :found]), {"found": found}
    return U, S, V, {"found": found}


def pm_iterations(ncols: int) -> int:
    """src/PM.cpp:25-28."""
    return int(_lib().oc_pm_iterations(ctypes.c_long(ncols)))


def rsvd(A, Omega, l: int, q: int = 2, method: int = JACOBI, seed: int = 0):
    """src/rSVD.cpp:72-133 with Omega supplied.  Returns (U, S, V) with the reference's output shapes."""
    A = np.asarray(A, dtype=np.float64)
    Q = intermediate_step(A, Omega, l, q)       # :84-85
    B = Q.T @ A                                 # :89
    if method == JACOBI:
        Ut, S, V, _ = svd_jacobi(B)
    elif method == PARALLEL_JACOBI:
        Ut, S, V, _ = svd_parallel_jacobi(B)
    elif method == POWER:
        Ut, S, V, _ = svd_power(B, 0, seed)
    else:
        raise ValueError("Unsupported SVD method")   # std::invalid_argument, :122-123
    return Q @ Ut, S, V                          # :128


# ---------------------------------------------------------------------------------------------------------------
# Image::normalize / compress / reconstruct / deNormalize (image_compression/src/image_com.cpp)
# ---------------------------------------------------------------------------------------------------------------
def image_normalize(A):
    """image_com.cpp:251-264.  Returns (normalised image, min, max); unchanged when min >= max."""
    A = np.asarray(A, dtype=np.float64); lo, hi = float(A.min()), float(A.max())
    return ((A - lo) / (hi - lo) if lo < hi else A.copy()), lo, hi


def image_denormalize(A, lo, hi):
    """image_com.cpp:270-281."""
    return A * (hi - lo) + lo if lo < hi else np.array(A, dtype=np.float64)


def image_compress(A, k: int = -1, Omega=None, seed: int = 0):
    """image_com.cpp:288-317: l = k + 10, older-API rSVD (image_compression/src/rSVD.cpp:77-118: q = 1, power-method SVD of B).
    Returns (U m x l, S l, V n x l)."""
    A = np.asarray(A, dtype=np.float64); m, n = A.shape
    if k == -1:
        k = min(m, n) // 4
    l = k + 10
    if Omega is None:
        Omega = np.random.default_rng(seed).standard_normal((n, l))
    Q = intermediate_step(A, Omega, l, 1)                    # image_compression/src/rSVD.cpp:103-106 (q = 1)
    B = Q.T @ A                                             # :109
    Ut, S, Vrows, _ = svd_power(B, 0, seed, resize=False)   # singularValueDecomposition(B, S, Utilde, V, min_dim), :112-114
    k = min(l, n)
    return Q @ Ut[:, :k], S[:k], np.asfortranarray(Vrows[:k, :].T)   # :117; the older API returns V with the vectors in columns


def image_reconstruct(U, S, V):
    """image_com.cpp:184-190."""
    return (U * S) @ V.T


# ---------------------------------------------------------------------------------------------------------------
# POD wrappers (POD/ParametricDiffusion1D/src/POD.cpp)
# ---------------------------------------------------------------------------------------------------------------
def pod_perform_svd(A, r: int, svd_type: int, Omega=None, seed: int = 0):
    """perform_SVD, POD.cpp:42-114.  Returns (U, sigma, V) in the layouts the selected back-end produces."""
    if svd_type == 0:
        return svd_power(A, r, seed)[:3]
    if svd_type == 1:
        return svd_jacobi(A)[:3]
    if svd_type == 2:
        return svd_parallel_jacobi(A)[:3]
    if svd_type in (3, 4, 5):
        if Omega is None:
            Omega = np.random.default_rng(seed).standard_normal((A.shape[1], r))
        return rsvd(A, Omega, r, 2, {3: POWER, 4: JACOBI, 5: PARALLEL_JACOBI}[svd_type], seed)
    raise ValueError("The svd_type should be in [0,5]. Check 'svd_type' in the parameter file.")     # :87-91 (std::exit there)


def _spd_sqrt(X):
    w, v = np.linalg.eigh(X)
    return (v * np.sqrt(w)) @ v.T


def pod(variant: int, S, r: int, tol: float = 0.0, svd_type: int = 1, Xh=None, D=None, Omega=None, seed: int = 0):
    """variant 0 naive_POD (:116-134), 1 standard_POD (:136-224), 2 energy_POD (:226-336), 3 weight_POD (:338-461).
    Returns (W, sigma) as the reference's public members: sigma are the singular values of the CORRELATION matrix for
    variants 1-3 and the modes are divided by them (:164-166)."""
    S = np.asarray(S, dtype=np.float64); Nh, ns = S.shape
    if variant == 0:
        U, sigma, _ = pod_perform_svd(S, r, svd_type, Omega, seed)
        return U, sigma
    Sm = S
    if ns <= Nh:
        if variant == 1:
            C = S.T @ S                                                   # :152
        else:
            if variant == 3:
                Sm = S @ _spd_sqrt(D)                                     # :363-370
            C = (Sm.T @ Xh) @ Sm                                          # :250, :373
        U, sigma, V = pod_perform_svd(C, r, svd_type, Omega, seed)
        W = np.zeros((Nh, r))
        for i in range(r):
            W[:, i] = Sm @ V[:, i] / sigma[i]                             # :164-166
    else:
        if variant == 1:
            K = S @ S.T                                                   # :171
            U, sigma, V = pod_perform_svd(K, r, svd_type, Omega, seed)
            W = U                                                         # :190
        else:
            Xs = _spd_sqrt(Xh)                                            # :272-273
            K = (((Xs @ S) @ S.T) @ Xs) if variant == 2 else ((((Xs @ S) @ D) @ S.T) @ Xs)   # :279, :408
            U, sigma, V = pod_perform_svd(K, r, svd_type, Omega, seed)
            W = np.linalg.solve(Xs, U[:, :r])                             # CG to 1e-12, :296-304
    s2 = sigma[:r] ** 2                                                   # :203-219
    den = s2.sum(); N = 0; I = 0.0; num = 0.0
    while I < (1 - tol ** 2) and N < r:
        num += s2[N]; I = num / den; N += 1
    return np.asfortranarray(W[:, :N]), sigma


# ---------------------------------------------------------------------------------------------------------------
# Givens QR / naive GEMM (API-surface helpers)
# ---------------------------------------------------------------------------------------------------------------
class PCA:
    """PCA<method>(data, normalize) -- PCA/include/PCA_class.hpp.  The SVD is the Jacobi / ParallelJacobi restatement
    above; the front/back steps are the reference's arithmetic written in numpy."""

    def __init__(self, data, normalize: bool = False, method: int = JACOBI):
        data = np.array(data, dtype=np.float64)
        if data.shape[0] < 2 or data.shape[1] < 2:                       # assertDataValid, :50-54
            raise ValueError("PCA requires at least 2 rows and 2 columns.")
        self.rows = data.shape[0]
        self.mean = data.sum(axis=0) / data.shape[0]                     # :33  colwise().mean()
        c = data - self.mean                                             # :34
        self.stddev = None
        if normalize:
            self.stddev = np.sqrt((c * c).sum(axis=0) / (data.shape[0] - 1))   # :39
            c = c / self.stddev                                          # :40
        self.centered = c
        self.U, self.S, self.V, _ = (svd_jacobi if method == JACOBI else svd_parallel_jacobi)(c)   # :45-46

    def explainedVariance(self):                                         # :76-79
        return self.S / np.sqrt(self.rows - 1)

    def explainedVarianceRatio(self):                                    # :81-84
        v = self.explainedVariance()
        return (v * v / (self.rows - 1)) / ((v * v).sum() / (self.rows - 1))

    def scores(self):                                                    # :86-88
        return self.U * self.S

    def loadings(self):                                                  # :90-92
        return self.V

    def projectToPCA(self, data):                                        # :93-95
        return (np.asarray(data, dtype=np.float64) - self.mean) @ self.V

    def reconstructFromPCA(self, pc):                                    # :97-99
        return np.asarray(pc, dtype=np.float64) @ self.V.T + self.mean

    def checkOrthogonality(self):                                        # :147-151
        return float(np.linalg.norm(self.V.T @ self.V - np.eye(self.V.shape[1])))


def givens_qr(A, reduced: bool = True):
    """src/QR.cpp:22-80."""
    A = _f(A); m, n = A.shape
    Q = np.zeros((m, m), order="F"); R = np.zeros((m, n), order="F")
    _lib().oc_givens_qr_full(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Q), _p(R))
    if reduced:
        return _f(Q[:, :n]), _f(R[:n, :])       # :78-79
    return Q, R


def manual_matmul(A, B):
    """src/matrixOperations.cpp:7-28."""
    A = _f(A); B = _f(B)
    C = np.zeros((A.shape[0], B.shape[1]), order="F")
    rc = _lib().oc_manual_matmul(_p(A), ctypes.c_long(A.shape[0]), ctypes.c_long(A.shape[1]), _p(B),
                                 ctypes.c_long(B.shape[0]), ctypes.c_long(B.shape[1]), _p(C))
    if rc != 0:
        raise ValueError("Matrices dimensions are not compatible for manual matrix multiplication")
    return C


# ---------------------------------------------------------------------------------------------------------------
# Comparison helpers (the reference's own notion of "matches": python/compare_rSVD.py:27-36 is a sign-agnostic
# element-wise mean; the parity tests use the subspace-invariant quantities of SURVEY.md 8d)
# ---------------------------------------------------------------------------------------------------------------
def sigma_close(s, s_ref, rtol=1e-8, floor=1e-6):
    s = np.asarray(s); s_ref = np.asarray(s_ref)
    tol = rtol * np.maximum(s_ref, floor * s_ref[0])
    return bool(np.all(np.abs(s - s_ref) <= tol)), float(np.max(np.abs(s - s_ref) / np.maximum(s_ref, floor * s_ref[0])))


def subspace_sin_theta(U1, U2):
    """sin of the largest principal angle between range(U1) and range(U2) (orthonormal columns)."""
    M = U2 - U1 @ (U1.T @ U2)
    return float(np.linalg.norm(M, 2))


def reconstruction_error(A, U, S, V):
    """||A - U diag(S) V^T||_F  (tests/rSVD_test.cpp:77-84)."""
    return float(np.linalg.norm(A - (U * S) @ V.T))


# ---------------------------------------------------------------------------------------------------------------
# The reference's own first-party sources compiled over the Eigen/MPI stand-in (dev container only)
# ---------------------------------------------------------------------------------------------------------------
class RefLib:
    """ctypes view of oracle/_ref/libref_rsvd.so (see oracle/ref_driver.cpp)."""

    def __init__(self):
        so = _HERE / "_ref" / "libref_rsvd.so"
        if not so.exists():
            build()
        if not so.exists():
            raise FileNotFoundError("oracle/_ref/libref_rsvd.so is not built (needs /root/reference)")
        self.lib = ctypes.CDLL(str(so))

    @staticmethod
    def available() -> bool:
        return (_HERE / "_ref" / "libref_rsvd.so").exists() or Path("/root/reference/src").is_dir()

    def intermediate_step(self, A, Omega, l, q):
        A = _f(A); Omega = _f(Omega); m, n = A.shape
        Q = np.zeros((m, l), order="F")
        self.lib.ref_intermediate_step(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Omega), ctypes.c_int(l), ctypes.c_int(q), _p(Q))
        return Q

    def _unpack(self, U, S, V, dims):
        d = list(dims)
        U2 = np.array(U.ravel(order="F")[: d[0] * d[1]]).reshape((d[0], d[1]), order="F")
        V2 = np.array(V.ravel(order="F")[: d[3] * d[4]]).reshape((d[3], d[4]), order="F")
        return U2, S[: d[2]].copy(), V2

    def rsvd(self, A, Omega, l, method=JACOBI):
        A = _f(A); Omega = _f(Omega); m, n = A.shape; cap = max(l, n, m)
        U = np.zeros(m * cap); S = np.zeros(cap); V = np.zeros(n * cap); dims = (ctypes.c_long * 5)()
        rc = self.lib.ref_rsvd(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Omega), ctypes.c_int(l), ctypes.c_int(method), _p(U), _p(S), _p(V), dims)
        if rc != 0:
            raise ValueError("Unsupported SVD method")
        return self._unpack(U, S, V, dims)

    def svd(self, A, method=JACOBI, r=0):
        A = _f(A); m, n = A.shape; cap = max(m, n)
        U = np.zeros(m * cap); S = np.zeros(cap); V = np.zeros(n * cap); dims = (ctypes.c_long * 5)()
        rc = self.lib.ref_svd(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(method), ctypes.c_int(r), _p(U), _p(S), _p(V), dims)
        if rc != 0:
            raise ValueError("Unsupported SVD method")
        return self._unpack(U, S, V, dims)

    def qr_reduced(self, A):
        A = _f(A); m, n = A.shape
        Q = np.zeros((m, n), order="F"); R = np.zeros((n, n), order="F")
        self.lib.ref_qr_reduced(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Q), _p(R))
        return Q, R

    def qr_full(self, A):
        A = _f(A); m, n = A.shape
        Q = np.zeros((m, m), order="F"); R = np.zeros((m, n), order="F")
        self.lib.ref_qr_full(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Q), _p(R))
        return Q, R

    def manual_matmul(self, A, B):
        A = _f(A); B = _f(B)
        C = np.zeros((A.shape[0], B.shape[1]), order="F")
        rc = self.lib.ref_manual_matmul(_p(A), ctypes.c_long(A.shape[0]), ctypes.c_long(A.shape[1]), _p(B), ctypes.c_long(B.shape[0]), ctypes.c_long(B.shape[1]), _p(C))
        if rc != 0:
            raise ValueError("Matrices dimensions are not compatible for manual matrix multiplication")
        return C

    def make_jacobi(self, x, y, z):
        c = ctypes.c_double(); s = ctypes.c_double()
        ok = self.lib.ref_make_jacobi(ctypes.c_double(x), ctypes.c_double(y), ctypes.c_double(z), ctypes.byref(c), ctypes.byref(s))
        return bool(ok), c.value, s.value

    def real_2x2_jacobi_svd(self, M):
        M = _f(M); out = [ctypes.c_double() for _ in range(4)]
        self.lib.ref_real_2x2_jacobi_svd(_p(M), *[ctypes.byref(o) for o in out])
        return tuple(o.value for o in out)

    def pm(self, A):
        A = _f(A); m, n = A.shape
        sigma = ctypes.c_double(); u = np.zeros(m); v = np.zeros(n)
        self.lib.ref_pm(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.byref(sigma), _p(u), _p(v))
        return sigma.value, u, v

    def pca(self, data, normalize=False, method=JACOBI, project=None):
        """PCA<method> through the reference's own class (oracle/ref_driver.cpp ref_pca).  Returns a dict."""
        data = _f(data); m, n = data.shape; k = min(m, n)
        P = _f(data if project is None else project); pr = P.shape[0]
        ev = np.zeros(k); ratio = np.zeros(k); scores = np.zeros((m, k), order="F"); load = np.zeros((n, k), order="F")
        mean = np.zeros(n); proj = np.zeros((pr, k), order="F"); recon = np.zeros((pr, n), order="F"); orth = ctypes.c_double(0)
        rc = self.lib.ref_pca(_p(data), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(int(normalize)), ctypes.c_int(method), _p(P),
                              ctypes.c_long(pr), _p(ev), _p(ratio), _p(scores), _p(load), _p(mean), _p(proj), _p(recon), ctypes.byref(orth))
        if rc == -1:
            raise ValueError("PCA requires at least 2 rows and 2 columns.")
        if rc != 0:
            raise ValueError("Unsupported SVD method")
        return dict(explained_variance=ev, ratio=ratio, scores=scores, loadings=load, mean=mean, project=proj, reconstruct=recon,
                    orthogonality=orth.value)

    def pod(self, variant, S, r, tol=0.0, svd_type=1, Xh=None, D=None, Omega=None):
        """The reference's POD class (oracle/ref_driver.cpp ref_pod).  Returns (W, sigma)."""
        S = _f(S); Nh, ns = S.shape; cap = max(Nh, ns)
        W = np.zeros(Nh * cap); sigma = np.zeros(cap); dims = (ctypes.c_long * 3)()
        Xh = _f(Xh) if Xh is not None else None; D = _f(D) if D is not None else None
        Om = _f(Omega) if Omega is not None else None
        rc = self.lib.ref_pod(ctypes.c_int(variant), _p(S), ctypes.c_long(Nh), ctypes.c_long(ns), _p(Xh) if Xh is not None else None,
                              _p(D) if D is not None else None, ctypes.c_int(r), ctypes.c_double(tol), ctypes.c_int(svd_type),
                              _p(Om) if Om is not None else None, ctypes.c_long(Om.shape[0] if Om is not None else 0), _p(W), _p(sigma), dims)
        if rc != 0:
            raise ValueError("bad POD variant")
        return np.array(W[: dims[0] * dims[1]]).reshape((dims[0], dims[1]), order="F"), sigma[: dims[2]].copy()


class RefLibV1:
    """ctypes view of oracle/_ref/libref_imgcomp.so: the reference's OLDER API (image_compression/src/*.cpp) and its Image class,
    compiled from the reference's own sources (oracle/ref_driver_v1.cpp)."""

    def __init__(self):
        so = _HERE / "_ref" / "libref_imgcomp.so"
        if not so.exists():
            build()
        if not so.exists():
            raise FileNotFoundError("oracle/_ref/libref_imgcomp.so is not built (needs /root/reference)")
        self.lib = ctypes.CDLL(str(so))

    @staticmethod
    def available() -> bool:
        return (_HERE / "_ref" / "libref_imgcomp.so").exists() or Path("/root/reference/image_compression/src").is_dir()

    def intermediate_step(self, A, Omega, l, q=1):
        A = _f(A); Omega = _f(Omega); m, n = A.shape
        Q = np.zeros((m, l), order="F")
        self.lib.ref1_intermediate_step(_p(A), ctypes.c_long(m), ctypes.c_long(n), _p(Omega), ctypes.c_int(l), ctypes.c_int(q), _p(Q))
        return Q

    def rsvd(self, A, l):
        """rSVD(A, U, S, V, l): Omega is drawn inside from std::random_device -- not reproducible."""
        A = _f(A); m, n = A.shape
        U = np.zeros((m, l), order="F"); S = np.zeros(l); V = np.zeros(n * max(l, n)); dims = (ctypes.c_long * 5)()
        self.lib.ref1_rsvd(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(l), _p(U), _p(S), _p(V), dims)
        return U, S, np.array(V[: dims[3] * dims[4]]).reshape((dims[3], dims[4]), order="F")

    def svd(self, A, dim):
        A = _f(A); m, n = A.shape
        S = np.zeros(dim); U = np.zeros((m, dim), order="F"); V = np.zeros((n, dim), order="F")
        self.lib.ref1_svd(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(dim), _p(S), _p(U), _p(V))
        return U, S, V

    def power_method(self, A):
        A = _f(A); m, n = A.shape
        sigma = ctypes.c_double(); u = np.zeros(m); v = np.zeros(n)
        self.lib.ref1_power_method(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.byref(sigma), _p(u), _p(v))
        return sigma.value, u, v

    def qr(self, A, reduced=True):
        A = _f(A); m, n = A.shape
        Q = np.zeros((m, n if reduced else m), order="F"); R = np.zeros((n if reduced else m, n), order="F")
        self.lib.ref1_qr(_p(A), ctypes.c_long(m), ctypes.c_long(n), ctypes.c_int(int(reduced)), _p(Q), _p(R))
        return Q, R

    def image_flow(self, path, rows, cols, scale, k):
        """Image: load -> downscale(scale) -> normalize -> compress(k) -> reconstruct.  rows x cols: the matrix shape the class
        will hold (width/scale x height/scale).  Returns dict(norm, lo, hi, S, recon, ratio, l)."""
        l = k + 10
        norm = np.zeros((rows, cols), order="F"); recon = np.zeros((rows, cols), order="F"); S = np.zeros(l)
        lo = ctypes.c_double(); hi = ctypes.c_double(); ratio = ctypes.c_double(); dims = (ctypes.c_long * 3)()
        rc = self.lib.ref1_image_flow(str(path).encode(), ctypes.c_int(scale), ctypes.c_int(k), _p(norm), ctypes.byref(lo), ctypes.byref(hi),
                                      _p(S), _p(recon), ctypes.byref(ratio), dims)
        if rc != 0:
            raise IOError(f"Image::load failed for {path}")
        assert (dims[0], dims[1], dims[2]) == (rows, cols, l), tuple(dims)
        return dict(norm=norm, lo=lo.value, hi=hi.value, S=S, recon=recon, ratio=ratio.value, l=l)
