"""Host-side mirror of the reference's rSVD API, in Python, over the C ABI of librsvdb.so.

The reference is C++ (Eigen types); its maintainers would use the C++ drop-in headers under include/.  This module is
the same surface for Python callers and for the parity tests, with the reference's names and argument meaning:

    rSVD(A, l, method)                 include/rSVD.hpp:14      -> (U, S, V)
    intermediate_step(A, Omega, l, q)  include/rSVD.hpp:13      -> Q
    generateOmega(n, l)                include/rSVD.hpp:15      -> Omega
    SVD(method)(A, r).compute()        include/SVD_class.hpp:35 -> getU / getS / getV
    qr_decomposition_reduced / _full   include/QR.hpp:15-16     -> (Q, R)
    PM(A)                              include/PM.hpp:18        -> (sigma, u, v)
    manualMatrixMultiply(A, B)         include/matrixOperations.hpp:14
    PCA(method)(data, normalize)       PCA/include/PCA_class.hpp:11 -> scores / loadings / explainedVariance / projectToPCA ...
    POD(S, [Xh, [D,]] r, [tol,] svd_type)   POD/ParametricDiffusion1D/src/POD.hpp:24 -> .W, .sigma

Host arrays are numpy float64 (any layout; they are passed column-major like Eigen::MatrixXd).  Every factorisation and
product runs in the CUDA library and there is no fallback; the only host arithmetic is the O(k) / O(m k) getter epilogues
of PCA (S / sqrt(m-1), U * diag(S)), which the reference also evaluates lazily on returned factors.
"""
from __future__ import annotations

import ctypes
from enum import IntEnum

import numpy as np

from . import capi


class SVDMethod(IntEnum):
    """enum class SVDMethod -- include/SVD_class.hpp:28-32."""
    Jacobi = 0
    Power = 1
    ParallelJacobi = 2


def _f(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Engine:
    """One librsvdb context (one GPU)."""

    def __init__(self, device: int = 0):
        self.lib = capi.load()
        h = ctypes.c_void_p()
        rc = self.lib.rsvdb_create(ctypes.byref(h), device)
        if rc != 0:
            raise capi.RsvdbError(rc, "rsvdb_create failed (is a B200 / sm_100 GPU visible?)")
        self.h = h
        self.device = device

    # -- plumbing -------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.rsvdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.rsvdb_last_error(self.h).decode()
            if rc == capi.ERR_INVALID_ARGUMENT:
                raise ValueError(msg)          # the reference throws std::invalid_argument
            raise capi.RsvdbError(rc, msg)

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.rsvdb_set_stream(self.h, ctypes.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.rsvdb_synchronize(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.rsvdb_launch_count(self.h))

    def set_qr_policy(self, householder_only: bool):
        """False (default): guarded CholeskyQR2 with Householder TSQR as the fallback; True: Householder TSQR always."""
        self._check(self.lib.rsvdb_set_qr_policy(self.h, 1 if householder_only else 0))

    def qr_path_counts(self):
        """(sketches orthonormalised by CholeskyQR2, by Householder TSQR) since the engine was created."""
        a = ctypes.c_int64(); b = ctypes.c_int64()
        self._check(self.lib.rsvdb_qr_path_counts(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def set_profiling(self, on: bool):
        self._check(self.lib.rsvdb_set_profiling(self.h, int(on)))

    def phase_ms(self) -> dict:
        out = (ctypes.c_double * 7)()
        self._check(self.lib.rsvdb_phase_ms(self.h, out))
        names = ["gemm_an", "gemm_at", "qr", "small_svd", "comm", "other", "copy"]
        return dict(zip(names, [float(x) for x in out]))

    def last_svd_info(self):
        s = ctypes.c_int(); r = ctypes.c_int()
        self._check(self.lib.rsvdb_last_svd_info(self.h, ctypes.byref(s), ctypes.byref(r)))
        return s.value, r.value

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = ctypes.create_string_buffer(uid, 128)
        self._check(self.lib.rsvdb_comm_init(self.h, nranks, rank, ctypes.cast(buf, ctypes.c_void_p)))

    def comm_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.rsvdb_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p))
        if rc != 0:
            raise capi.RsvdbError(rc, "ncclGetUniqueId failed")
        return buf.raw

    # -- device-pointer entry points (raw addresses; the caller owns the memory, e.g. torch tensors) ------------------------
    def gemm_an_dev(self, dA, m, n, lda, dX, ldx, l, dY, ldy):
        self._check(self.lib.rsvdb_gemm_an_dev(self.h, dA, m, n, lda, dX, ldx, l, dY, ldy))

    def gemm_at_dev(self, dA, m, n, lda, dQ, ldq, l, dZ, ldz, transpose_out=False):
        self._check(self.lib.rsvdb_gemm_at_dev(self.h, dA, m, n, lda, dQ, ldq, l, dZ, ldz, int(transpose_out)))

    def qr_dev(self, dY, rows, l, ldy, sharded=False, dR=None):
        self._check(self.lib.rsvdb_qr_dev(self.h, dY, rows, l, ldy, int(sharded), dR))

    def orthonormalize_dev(self, dY, rows, l, ldy, sharded=False, dR=None) -> int:
        """Y <- orthonormal basis of its columns as the rSVD pipeline computes it; returns 0 (CholeskyQR2) or 1 (Householder)."""
        path = ctypes.c_int(-1)
        self._check(self.lib.rsvdb_orthonormalize_dev(self.h, dY, rows, l, ldy, int(sharded), dR, ctypes.byref(path)))
        return path.value

    def range_finder_dev(self, dA, m_local, n, lda, dOmega, ldo, l, q, dQ, ldq):
        self._check(self.lib.rsvdb_range_finder_dev(self.h, dA, m_local, n, lda, dOmega, ldo, l, q, dQ, ldq))

    def rsvd_dev(self, dA, m_local, n, lda, dOmega, ldo, l, q, method, dU, ldu, dS, dV, ldv, seed=0):
        self._check(self.lib.rsvdb_rsvd_dev(self.h, dA, m_local, n, lda, dOmega, ldo, l, q, int(method), seed, dU, ldu, dS, dV, ldv))

    def generate_omega_dev(self, n, l, seed, dOmega, ldo):
        self._check(self.lib.rsvdb_generate_omega_dev(self.h, n, l, seed, dOmega, ldo))

    def rsvd_host_raw(self, pA, m, n, lda, pOmega, ldo, seed, l, q, method, pU, ldu, pS, pV, ldv):
        """rsvdb_rsvd_host on raw host addresses (e.g. pinned torch tensors)."""
        self._check(self.lib.rsvdb_rsvd_host(self.h, pA, m, n, lda, pOmega, ldo, seed, l, q, int(method), pU, ldu, pS, pV, ldv))

    # -- reference API mirrors (host arrays) --------------------------------------------------------------------------
    def generateOmega(self, n: int, l: int, seed: int = 0) -> np.ndarray:
        """src/rSVD.cpp:12-55 (N(0,1) entries; seeded here, std::random_device there)."""
        Om = np.zeros((n, l), order="F")
        self._check(self.lib.rsvdb_generate_omega_host(self.h, n, l, seed, _ptr(Om), max(n, 1)))
        return Om

    def intermediate_step(self, A, Omega, l: int, q: int) -> np.ndarray:
        """src/rSVD.cpp:57-70."""
        A = _f(A); Omega = _f(Omega); m, n = A.shape
        if Omega.shape != (n, l):
            raise ValueError("Omega must be n x l")
        Q = np.zeros((m, l), order="F")
        self._check(self.lib.rsvdb_intermediate_step_host(self.h, _ptr(A), m, n, max(m, 1), _ptr(Omega), n, l, q, _ptr(Q), max(m, 1)))
        return Q

    def rSVD(self, A, l: int, method=SVDMethod.Jacobi, Omega=None, q: int = 2, seed: int = 0):
        """rSVD(A, U, S, V, l, method) -- src/rSVD.cpp:72-133.  Returns (U, S, V) with the reference's output shapes:
        Jacobi / ParallelJacobi: U m x k, S k, V n x k;  Power: U m x l, S l, V n x n with the vectors in ROWS."""
        A = _f(A); m, n = A.shape
        try:
            method = int(method)
        except Exception as e:
            raise ValueError("Unsupported SVD method") from e
        k = min(l, n)
        U = np.zeros((m, max(k, 1)), order="F"); S = np.zeros(max(k, 1)); V = np.zeros((n, max(k, 1)), order="F")
        om_ptr, ldo = None, 0
        if Omega is not None:
            Omega = _f(Omega)
            if Omega.shape != (n, l):
                raise ValueError("Omega must be n x l")
            om_ptr, ldo = _ptr(Omega), n
        self._check(self.lib.rsvdb_rsvd_host(self.h, _ptr(A), m, n, max(m, 1), om_ptr, ldo, seed, l, q, method,
                                             _ptr(U), max(m, 1), _ptr(S), _ptr(V), n))
        if method == SVDMethod.Power:
            return U, S, _power_v_layout(V, n, k, k)
        return U, S, V

    def rSVD_csr(self, rowptr, colidx, values, shape, l: int, method=SVDMethod.Jacobi, Omega=None, q: int = 2, seed: int = 0):
        """rSVD of a CSR matrix (int64 row pointers, int32 column indices) without densifying it -- the reference
        densifies every .mtx first (tests/rSVD_test.cpp:54-57).  Same outputs as rSVD()."""
        m, n = shape
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64); colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        nnz = int(rowptr[-1])
        k = min(l, n)
        U = np.zeros((m, max(k, 1)), order="F"); S = np.zeros(max(k, 1)); V = np.zeros((n, max(k, 1)), order="F")
        om_ptr, ldo = None, 0
        if Omega is not None:
            Omega = _f(Omega); om_ptr, ldo = _ptr(Omega), n
        self._check(self.lib.rsvdb_rsvd_csr_host(self.h, m, n, nnz, rowptr.ctypes.data, colidx.ctypes.data, values.ctypes.data, om_ptr, ldo,
                                                 seed, l, q, int(method), _ptr(U), max(m, 1), _ptr(S), _ptr(V), n))
        if int(method) == SVDMethod.Power:
            return U, S, _power_v_layout(V, n, k, k)
        return U, S, V

    def qr_decomposition_reduced(self, A):
        """src/QR.cpp:43-80: Q m x n, R n x n."""
        A = _f(A); m, n = A.shape
        Q = np.zeros((m, n), order="F"); R = np.zeros((n, n), order="F")
        self._check(self.lib.rsvdb_qr_host(self.h, _ptr(A), m, n, m, 0, _ptr(Q), m, _ptr(R), n))
        return Q, R

    def qr_decomposition_full(self, A):
        """src/QR.cpp:22-41: Q m x m, R m x n."""
        A = _f(A); m, n = A.shape
        Q = np.zeros((m, m), order="F"); R = np.zeros((m, n), order="F")
        self._check(self.lib.rsvdb_qr_host(self.h, _ptr(A), m, n, m, 1, _ptr(Q), m, _ptr(R), m))
        return Q, R

    def PM(self, A, seed: int = 0):
        """PM(A, B, sigma, u, v) -- src/PM.cpp:4-81."""
        A = _f(A); m, n = A.shape
        u = np.zeros(m); v = np.zeros(n); sigma = ctypes.c_double()
        self._check(self.lib.rsvdb_pm_host(self.h, _ptr(A), m, n, m, seed, ctypes.byref(sigma), _ptr(u), _ptr(v)))
        return sigma.value, u, v

    def manualMatrixMultiply(self, A, B):
        """src/matrixOperations.cpp:7-28."""
        A = _f(A); B = _f(B)
        C = np.zeros((A.shape[0], B.shape[1]), order="F")
        self._check(self.lib.rsvdb_gemm_host(self.h, _ptr(A), A.shape[0], A.shape[1], max(A.shape[0], 1), _ptr(B), B.shape[0],
                                             B.shape[1], max(B.shape[0], 1), _ptr(C), max(A.shape[0], 1)))
        return C

    def svd(self, A, method=SVDMethod.Jacobi, r: int = 0, seed: int = 0):
        """SVD<method>(A, r).compute() -- include/SVD_class.hpp:79-97.  Returns (U, S, V) in the reference's shapes."""
        A = _f(A); m, n = A.shape; k = min(m, n)
        try:
            method = int(method)
        except Exception as e:
            raise ValueError("Unsupported SVD method") from e
        found = ctypes.c_int()
        if method == SVDMethod.Power:
            dim = r if r else k
            U = np.zeros((m, m), order="F"); S = np.zeros(k); V = np.zeros((n, max(dim, 1)), order="F")
            self._check(self.lib.rsvdb_svd_host(self.h, _ptr(A), m, n, m, method, r, seed, _ptr(U), m, _ptr(S), _ptr(V), n, ctypes.byref(found)))
            f = found.value
            Vref = _power_v_layout(V, n, dim, f)
            if f < dim:   # conservativeResize on the early exit, include/SVD_class.hpp:198-209
                if f == 0:
                    return np.zeros((m, 1), order="F"), np.zeros(1), np.zeros((n, 1), order="F")
                return _f(U[:, :f]), S[:f].copy(), _f(Vref[:, :f])
            return U, S, Vref
        U = np.zeros((m, k), order="F"); S = np.zeros(k); V = np.zeros((n, k), order="F")
        self._check(self.lib.rsvdb_svd_host(self.h, _ptr(A), m, n, m, method, r, seed, _ptr(U), m, _ptr(S), _ptr(V), n, ctypes.byref(found)))
        return U, S, V


    # -- PCA front / back steps (PCA/include/PCA_class.hpp) ------------------------------------------------------------
    def pca(self, data, normalize: bool = False, method=SVDMethod.Jacobi, r: int = 0, seed: int = 0):
        """PCA<method>(data, normalize)::initialize() -- PCA_class.hpp:24-47: centre, optionally scale, SVD<method>; one
        upload.  Returns (mean, stddev or None, U, S, V)."""
        A = _f(data); m, n = A.shape; k = min(m, n)
        method = int(method)
        if method == SVDMethod.Power:
            raise ValueError("PCA<Power>: scores() is ill-formed in the reference (U is m x m); use Jacobi / ParallelJacobi")
        mean = np.zeros(n); sd = np.zeros(n)
        U = np.zeros((m, max(k, 1)), order="F"); S = np.zeros(max(k, 1)); V = np.zeros((n, max(k, 1)), order="F")
        found = ctypes.c_int()
        self._check(self.lib.rsvdb_pca_host(self.h, _ptr(A), m, n, max(m, 1), int(bool(normalize)), method, r, seed, _ptr(mean), _ptr(sd),
                                            _ptr(U), max(m, 1), _ptr(S), _ptr(V), max(n, 1), ctypes.byref(found)))
        return mean, (sd if normalize else None), U, S, V

    def rpca(self, data, l: int, normalize: bool = False, method=SVDMethod.Jacobi, Omega=None, q: int = 2, seed: int = 0):
        """Randomized PCA: rSVD of the centred (scaled) matrix without materialising it.  Returns (mean, stddev, U, S, V)."""
        A = _f(data); m, n = A.shape; k = min(l, n)
        mean = np.zeros(n); sd = np.zeros(n)
        U = np.zeros((m, k), order="F"); S = np.zeros(k); V = np.zeros((n, k), order="F")
        om_ptr, ldo = None, 0
        if Omega is not None:
            Omega = _f(Omega)
            if Omega.shape != (n, l):
                raise ValueError("Omega must be n x l")
            om_ptr, ldo = _ptr(Omega), n
        self._check(self.lib.rsvdb_rpca_host(self.h, _ptr(A), m, n, max(m, 1), int(bool(normalize)), om_ptr, ldo, seed, l, q, int(method),
                                             _ptr(mean), _ptr(sd), _ptr(U), max(m, 1), _ptr(S), _ptr(V), n))
        return mean, (sd if normalize else None), U, S, V

    def pca_project(self, data, mean, V):
        """projectToPCA -- PCA_class.hpp:93-95."""
        D = _f(data); V = _f(V); r, n = D.shape; k = V.shape[1]
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        out = np.zeros((r, k), order="F")
        self._check(self.lib.rsvdb_pca_project_host(self.h, _ptr(D), r, n, r, _ptr(mean), _ptr(V), n, k, _ptr(out), r))
        return out

    def pca_reconstruct(self, pc, mean, V):
        """reconstructFromPCA -- PCA_class.hpp:97-99."""
        P = _f(pc); V = _f(V); r, k = P.shape; n = V.shape[0]
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        out = np.zeros((r, n), order="F")
        self._check(self.lib.rsvdb_pca_reconstruct_host(self.h, _ptr(P), r, k, r, _ptr(mean), _ptr(V), n, n, _ptr(out), r))
        return out


    # -- POD wrappers (POD/ParametricDiffusion1D/src/POD.cpp) -----------------------------------------------------------
    def pod(self, variant: int, S, r: int, tol: float = 0.0, svd_type: int = 1, Xh=None, D=None, Omega=None, seed: int = 0):
        """naive (0) / standard (1) / energy (2) / weight (3) POD -- POD.cpp:116-461; svd_type as perform_SVD (:42-114).
        Returns (W, sigma): the truncated basis and the singular values exactly as the reference stores them."""
        S = _f(S); Nh, ns = S.shape
        wc = ctypes.c_int64(); sl = ctypes.c_int64()
        if self.lib.rsvdb_pod_shape(int(variant), Nh, ns, int(r), int(svd_type), ctypes.byref(wc), ctypes.byref(sl)) != 0:
            if not 0 <= int(svd_type) <= 5:
                raise ValueError("The svd_type should be in [0,5]. Check 'svd_type' in the parameter file.")
            raise ValueError("POD: bad argument")
        W = np.zeros((Nh, wc.value), order="F"); sigma = np.zeros(sl.value); N = ctypes.c_int()
        Xh = _f(Xh) if Xh is not None else None; D = _f(D) if D is not None else None
        om_ptr, ldo = None, 0
        if Omega is not None:
            Omega = _f(Omega); om_ptr, ldo = _ptr(Omega), Omega.shape[0]
        self._check(self.lib.rsvdb_pod_host(self.h, int(variant), _ptr(S), Nh, ns, Nh, _ptr(Xh) if Xh is not None else None, Nh,
                                            _ptr(D) if D is not None else None, ns, int(r), float(tol), int(svd_type), seed, om_ptr, ldo,
                                            _ptr(W), Nh, _ptr(sigma), ctypes.byref(N)))
        return np.asfortranarray(W[:, :N.value]), sigma


    # -- Image::compress-shaped driver (image_compression/src/image_com.cpp) --------------------------------------------
    def image_compress(self, image, k: int = -1, normalize: bool = True, Omega=None, seed: int = 0):
        """Image::normalize + Image::compress in one upload -- image_com.cpp:251-264,288-317.  Returns (U, S, V, min, max, l)."""
        A = _f(image); m, n = A.shape
        kk = min(m, n) // 4 if k == -1 else k
        l = kk + 10
        if l > n or kk < 0:
            raise ValueError("Image::compress: k + 10 must not exceed the image width")
        U = np.zeros((m, l), order="F"); S = np.zeros(l); V = np.zeros((n, l), order="F")
        lo = ctypes.c_double(); hi = ctypes.c_double(); deg = ctypes.c_int()
        om_ptr, ldo = None, 0
        if Omega is not None:
            Omega = _f(Omega); om_ptr, ldo = _ptr(Omega), n
        self._check(self.lib.rsvdb_image_compress_host(self.h, _ptr(A), m, n, m, int(k), int(bool(normalize)), om_ptr, ldo, seed, ctypes.byref(lo),
                                                       ctypes.byref(hi), _ptr(U), m, _ptr(S), _ptr(V), n, ctypes.byref(deg)))
        return U, S, V, lo.value, hi.value, deg.value

    def image_normalize(self, image, inverse: bool = False, lo: float = 0.0, hi: float = 0.0):
        """Image::normalize / deNormalize -- image_com.cpp:251-281.  Returns (mapped image, min, max)."""
        A = _f(image).copy(order="F"); m, n = A.shape
        clo = ctypes.c_double(lo); chi = ctypes.c_double(hi)
        self._check(self.lib.rsvdb_image_normalize_host(self.h, _ptr(A), m, n, m, int(bool(inverse)), ctypes.byref(clo), ctypes.byref(chi)))
        return A, clo.value, chi.value

    def image_reconstruct(self, U, S, V, denormalize: bool = False, lo: float = 0.0, hi: float = 1.0):
        """Image::reconstruct (+ deNormalize in the same pass) -- image_com.cpp:184-190,270-281."""
        U = _f(U); V = _f(V); S = np.ascontiguousarray(S, dtype=np.float64); m, l = U.shape; n = V.shape[0]
        out = np.zeros((m, n), order="F")
        self._check(self.lib.rsvdb_image_reconstruct_host(self.h, _ptr(U), m, m, _ptr(S), _ptr(V), n, n, l, int(bool(denormalize)), float(lo), float(hi),
                                                          _ptr(out), m))
        return out


def _power_v_layout(Vcols: np.ndarray, n: int, dim: int, found: int) -> np.ndarray:
    """The Power back-end stores right singular vectors in the ROWS of an identity-initialised n x n matrix
    (include/SVD_class.hpp:83,214)."""
    V = np.asfortranarray(np.eye(n))
    for i in range(min(found, dim)):
        V[i, :] = Vcols[:, i]
    return V


class SVD:
    """template<SVDMethod> class SVD -- include/SVD_class.hpp:35-71."""

    def __init__(self, engine: Engine, method, data, r: int = 0, seed: int = 0):
        self._e, self._method, self._r, self._seed = engine, method, r, seed
        self._data = _f(data).copy(order="F")        # the reference copies its input (data_, :55,74-75)
        self._U = self._S = self._V = None

    def setData(self, data):                           # protected in the reference (:67-70), used by PCA
        self._data = _f(data).copy(order="F")

    def compute(self):
        self._U, self._S, self._V = self._e.svd(self._data, self._method, self._r, self._seed)

    def getU(self): return self._U.copy(order="F")     # getters return by value (:46-48)
    def getS(self): return self._S.copy()
    def getV(self): return self._V.copy(order="F")


class PCA(SVD):
    """template<SVDMethod> class PCA : public SVD<method> -- PCA/include/PCA_class.hpp:11-207 (same member names).
    ``l`` (additive): when given, the decomposition is the rank-l randomized one of ``Engine.rpca`` instead of the full SVD."""

    def __init__(self, engine: Engine, method, data, normalize: bool = False, l: int = 0, Omega=None, q: int = 2):
        super().__init__(engine, method, data)
        self._orig = _f(data).copy(order="F")
        self._normalize = bool(normalize)
        self._l, self._Omega, self._q = l, Omega, q
        self.initialize()

    def initialize(self):                                   # :24-47
        self.assertDataValid()
        if self._l:
            self._mean, self._stddev, self._U, self._S, self._V = self._e.rpca(self._orig, self._l, self._normalize, self._method,
                                                                                 self._Omega, self._q)
        else:
            self._mean, self._stddev, self._U, self._S, self._V = self._e.pca(self._orig, self._normalize, self._method)

    def assertDataValid(self):                              # :50-54
        if self._orig.shape[0] < 2 or self._orig.shape[1] < 2:
            raise ValueError("PCA requires at least 2 rows and 2 columns.")

    def addData(self, new_data):                            # :57-61
        self._orig = np.asfortranarray(np.vstack([self._orig, _f(new_data)]))
        self.initialize()

    def normalizeData(self):                                # :63-66 (scales the stored data by the UNcentred second moment)
        sd = np.sqrt((self._orig ** 2).sum(axis=0) / (self._orig.shape[0] - 1))
        self._stddev = sd
        self._orig = np.asfortranarray(self._orig / sd)

    def setNormalization(self, normalize: bool):            # :69-72
        self._normalize = bool(normalize)
        self.initialize()

    def explainedVariance(self):                            # :75-78
        return self.getS() / np.sqrt(self._orig.shape[0] - 1)

    def explainedVarianceRatio(self):                       # :80-83
        v = self.explainedVariance(); d = self._orig.shape[0] - 1
        return (v * v / d) / ((v * v).sum() / d)

    def scores(self):                                       # :85-87
        return np.asfortranarray(self.getU() * self.getS())

    def loadings(self):                                     # :89-91
        return self.getV()

    def projectToPCA(self, data):                           # :93-95
        return self._e.pca_project(data, self._mean, self._V)

    def reconstructFromPCA(self, pc):                       # :97-99
        return self._e.pca_reconstruct(pc, self._mean, self._V)

    def checkOrthogonality(self) -> float:                  # :147-151
        V = self.getV()
        return float(np.linalg.norm(V.T @ V - np.eye(V.shape[1])))

    def mean(self): return self._mean.copy()                # additive: mean_ / stddev_ are private without getters in the reference
    def stddev(self): return None if self._stddev is None else self._stddev.copy()

    def summary(self) -> str:                               # :153-196 (returned instead of printed)
        ev = self.explainedVariance(); pr = self.explainedVarianceRatio(); cum = np.cumsum(pr)
        head = f"{'Component':<25}" + "".join(f"{'Comp.' + str(i + 1):<15}" for i in range(ev.size))
        rows = [("Standard deviation", ev), ("Proportion of Variance", pr), ("Cumulative Proportion", cum)]
        return "\n".join(["Importance of components:", head] + [f"{nm:<25}" + "".join(f"{x:<15.6f}" for x in v) for nm, v in rows])

    def saveResults(self, filename):                        # :101-144 (same file layout)
        cum = np.cumsum(self.explainedVarianceRatio())
        with open(filename, "w") as f:
            f.write("\nCumulative Explained Variance:\n")
            for x in cum:
                f.write(f"{x:g}\n")
            for title, M in (("Scores", self.scores()), ("Loadings", self.loadings())):
                f.write(f"\n{title}:\n")
                for row in M:
                    f.write(", ".join(f"{x:g}" for x in row) + "\n")


class Image:
    """class Image -- image_compression/include/image_comp.hpp:16-118, minus the stb codec (load / save), which is out of
    scope: the pixel matrix comes in through setMatrix.  normalize / compress / reconstruct / deNormalize run on the device."""

    def __init__(self, engine: Engine, width: int = 0, height: int = 0):
        self._e = engine
        self.originalWidth, self.originalHeight = width, height
        self.image_matrix = None
        self.left_singular = self.singular = self.right_singular = None
        self.original_min = self.original_max = 0.0
        self.degree = 0

    def setMatrix(self, M):                                   # additive (the reference fills image_matrix in load(), :18-44)
        self.image_matrix = _f(M).copy(order="F")
        self.originalHeight, self.originalWidth = self.image_matrix.shape[1], self.image_matrix.shape[0]   # load() stores the transpose

    def getMatrix(self):
        return self.image_matrix.copy(order="F")

    def normalize(self):                                      # :251-264
        self.image_matrix, self.original_min, self.original_max = self._e.image_normalize(self.image_matrix)

    def deNormalize(self):                                    # :270-281
        self.image_matrix, _, _ = self._e.image_normalize(self.image_matrix, True, self.original_min, self.original_max)

    def compress(self, k: int = -1, Omega=None, seed: int = 0):   # :288-317
        U, S, V, _, _, l = self._e.image_compress(self.image_matrix, k, False, Omega, seed)
        self.left_singular, self.singular, self.right_singular, self.degree = U, S, V, l

    def normalize_and_compress(self, k: int = -1, Omega=None, seed: int = 0):
        """additive: normalize() + compress(k) with ONE upload; image_matrix itself is left untouched"""
        U, S, V, self.original_min, self.original_max, l = self._e.image_compress(self.image_matrix, k, True, Omega, seed)
        self.left_singular, self.singular, self.right_singular, self.degree = U, S, V, l

    def reconstruct(self, denormalize: bool = False):         # :184-190 (denormalize: additive, fused deNormalize)
        return self._e.image_reconstruct(self.left_singular, self.singular, self.right_singular, denormalize, self.original_min, self.original_max)

    def downscale(self, scale_factor: int = -1):              # :193-217 (index shuffles on the host, like the reference)
        f = 2 if scale_factor == -1 else scale_factor
        nw, nh = self.originalWidth // f, self.originalHeight // f
        self.image_matrix = np.asfortranarray(self.image_matrix[: nh * f: f, : nw * f: f][:nh, :nw])
        self.originalWidth, self.originalHeight = nw, nh

    def upscale(self, scale_factor: int = -1):                # :219-244
        f = 2 if scale_factor == -1 else scale_factor
        self.image_matrix = np.asfortranarray(np.kron(self.image_matrix[: self.originalHeight, : self.originalWidth], np.ones((f, f))))
        self.originalWidth, self.originalHeight = self.originalWidth * f, self.originalHeight * f

    def get_compression_ratio(self) -> float:                 # :406-411
        return (self.originalHeight * self.originalWidth) / (self.degree * (self.originalWidth + self.originalHeight + 1))


class POD:
    """class POD -- POD/ParametricDiffusion1D/src/POD.hpp:24-70.  The four constructors are told apart by their argument
    lists like the C++ overloads: (S, r, svd_type) naive; (S, r, tol, svd_type) standard; (S, Xh, r, tol, svd_type) energy;
    (S, Xh, D, r, tol, svd_type) weight.  Public members W (POD modes) and sigma, as in the reference."""

    def __init__(self, engine: Engine, S, *args, Omega=None, seed: int = 0):
        if len(args) == 2:
            variant, Xh, D, (r, svd_type), tol = 0, None, None, args, 0.0
        elif len(args) == 3:
            variant, Xh, D, (r, tol, svd_type) = 1, None, None, args
        elif len(args) == 4:
            variant, D, (Xh, r, tol, svd_type) = 2, None, args
        elif len(args) == 5:
            variant, (Xh, D, r, tol, svd_type) = 3, args
        else:
            raise TypeError("POD(S, r, svd_type) | POD(S, r, tol, svd_type) | POD(S, Xh, r, tol, svd_type) | POD(S, Xh, D, r, tol, svd_type)")
        self.W, self.sigma = engine.pod(variant, S, r, tol, svd_type, Xh, D, Omega, seed)


_default = None


def default_engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine(0)
    return _default


# free functions with the reference's names
def rSVD(A, l, method=SVDMethod.Jacobi, Omega=None, q=2, seed=0):
    return default_engine().rSVD(A, l, method, Omega, q, seed)


def intermediate_step(A, Omega, l, q):
    return default_engine().intermediate_step(A, Omega, l, q)


def generateOmega(n, l, seed=0):
    return default_engine().generateOmega(n, l, seed)
