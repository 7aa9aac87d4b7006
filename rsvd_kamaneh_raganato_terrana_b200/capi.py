"""ctypes binding of librsvdb.so (the C ABI declared in include/rsvdb.h).

The library is the product: if it is missing or fails to load, this module raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p, POINTER
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "librsvdb.so"

OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NCCL, ERR_ALLOC, ERR_NO_CONVERGENCE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
SVD_JACOBI, SVD_POWER, SVD_PARALLEL_JACOBI = 0, 1, 2


class RsvdbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rsvdb error {code}: {msg}")
        self.code = code


_lib = None


def load() -> ctypes.CDLL:
    """Load librsvdb.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -m rsvd_kamaneh_raganato_terrana_b200.build` "
            "(nvcc, sm_100a). There is no fallback implementation."
        )
    lib = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib):
    vp = c_void_p  # device or host pointers are passed as raw addresses
    i64, u64 = c_int64, ctypes.c_uint64
    lib.rsvdb_version.restype = c_char_p
    lib.rsvdb_last_error.restype = c_char_p
    lib.rsvdb_launch_count.restype = c_int64
    lib.rsvdb_generic_gemm_fallbacks.restype = c_int64
    lib.rsvdb_split_gemm_products.restype = c_int64
    sig = {
        "rsvdb_create": [POINTER(c_void_p), c_int],
        "rsvdb_destroy": [vp],
        "rsvdb_set_stream": [vp, vp],
        "rsvdb_use_own_stream": [vp],
        "rsvdb_synchronize": [vp],
        "rsvdb_last_error": [vp],
        "rsvdb_launch_count": [vp],
        "rsvdb_generic_gemm_fallbacks": [],
        "rsvdb_split_gemm_products": [],
        "rsvdb_set_qr_policy": [vp, c_int],
        "rsvdb_qr_path_counts": [vp, POINTER(c_int64), POINTER(c_int64)],
        "rsvdb_set_profiling": [vp, c_int],
        "rsvdb_phase_ms": [vp, POINTER(c_double)],
        "rsvdb_last_svd_info": [vp, POINTER(c_int), POINTER(c_int)],
        "rsvdb_comm_unique_id": [vp],
        "rsvdb_comm_init": [vp, c_int, c_int, vp],
        "rsvdb_comm_size": [vp],
        "rsvdb_comm_rank": [vp],
        "rsvdb_pm_iterations": [i64],
        "rsvdb_gemm_an_dev": [vp, vp, i64, i64, i64, vp, i64, c_int, vp, i64],
        "rsvdb_gemm_at_dev": [vp, vp, i64, i64, i64, vp, i64, c_int, vp, i64, c_int],
        "rsvdb_qr_dev": [vp, vp, i64, c_int, i64, c_int, vp],
        "rsvdb_orthonormalize_dev": [vp, vp, i64, c_int, i64, c_int, vp, POINTER(c_int)],
        "rsvdb_range_finder_dev": [vp, vp, i64, i64, i64, vp, i64, c_int, c_int, vp, i64],
        "rsvdb_rsvd_dev": [vp, vp, i64, i64, i64, vp, i64, c_int, c_int, c_int, u64, vp, i64, vp, vp, i64],
        "rsvdb_generate_omega_dev": [vp, i64, c_int, u64, vp, i64],
        "rsvdb_rsvd_host": [vp, vp, i64, i64, i64, vp, i64, u64, c_int, c_int, c_int, vp, i64, vp, vp, i64],
        "rsvdb_intermediate_step_host": [vp, vp, i64, i64, i64, vp, i64, c_int, c_int, vp, i64],
        "rsvdb_generate_omega_host": [vp, i64, c_int, u64, vp, i64],
        "rsvdb_svd_host": [vp, vp, i64, i64, i64, c_int, c_int, u64, vp, i64, vp, vp, i64, POINTER(c_int)],
        "rsvdb_pca_host": [vp, vp, i64, i64, i64, c_int, c_int, c_int, u64, vp, vp, vp, i64, vp, vp, i64, POINTER(c_int)],
        "rsvdb_column_stats_dev": [vp, vp, i64, i64, i64, vp, vp],
        "rsvdb_center_columns_dev": [vp, vp, i64, i64, i64, vp, vp],
        "rsvdb_rpca_dev": [vp, vp, i64, i64, i64, vp, vp, vp, i64, c_int, c_int, c_int, vp, i64, vp, vp, i64],
        "rsvdb_rpca_host": [vp, vp, i64, i64, i64, c_int, vp, i64, u64, c_int, c_int, c_int, vp, vp, vp, i64, vp, vp, i64],
        "rsvdb_pca_project_host": [vp, vp, i64, i64, i64, vp, vp, i64, c_int, vp, i64],
        "rsvdb_pca_reconstruct_host": [vp, vp, i64, c_int, i64, vp, vp, i64, i64, vp, i64],
        "rsvdb_pod_shape": [c_int, i64, i64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)],
        "rsvdb_pod_host": [vp, c_int, vp, i64, i64, i64, vp, i64, vp, i64, c_int, c_double, c_int, u64, vp, i64, vp, i64, vp, POINTER(c_int)],
        "rsvdb_pod_dev": [vp, c_int, vp, i64, i64, i64, vp, i64, vp, i64, c_int, c_double, c_int, u64, vp, i64, vp, i64, vp, POINTER(c_int)],
        "rsvdb_image_compress_host": [vp, vp, i64, i64, i64, c_int, c_int, vp, i64, u64, POINTER(c_double), POINTER(c_double), vp, i64, vp, vp, i64,
                                      POINTER(c_int)],
        "rsvdb_image_normalize_host": [vp, vp, i64, i64, i64, c_int, POINTER(c_double), POINTER(c_double)],
        "rsvdb_image_reconstruct_host": [vp, vp, i64, i64, vp, vp, i64, i64, c_int, c_int, c_double, c_double, vp, i64],
        "rsvdb_qr_host": [vp, vp, i64, i64, i64, c_int, vp, i64, vp, i64],
        "rsvdb_pm_host": [vp, vp, i64, i64, i64, u64, POINTER(c_double), vp, vp],
        "rsvdb_gemm_host": [vp, vp, i64, i64, i64, vp, i64, i64, i64, vp, i64],
        "rsvdb_rsvd_csr_host": [vp, i64, i64, i64, vp, vp, vp, vp, i64, u64, c_int, c_int, c_int, vp, i64, vp, vp, i64],
        "rsvdb_rsvd_csr_dev": [vp, i64, i64, i64, vp, vp, vp, vp, i64, u64, c_int, c_int, c_int, vp, i64, vp, vp, i64],
        "rsvdb_csr_spmm_dev": [vp, i64, vp, vp, vp, vp, c_int, vp],
    }
    for name, argtypes in sig.items():
        getattr(lib, name).argtypes = argtypes


def exported_symbols() -> list[str]:
    """Symbols include/rsvdb.h declares (parsed from the header) -- used by the CPU test that the library exports them."""
    import re
    hdr = (_PKG.parent / "include" / "rsvdb.h").read_text()
    return sorted(set(re.findall(r"RSVDB_API\s+[\w\s\*]+?\b(rsvdb_\w+)\s*\(", hdr)))
