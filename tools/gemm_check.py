"""First GPU trip: correctness of the DMMA/TMA skinny GEMMs against torch.matmul(float64) and timing at C5-like shapes."""
import ctypes, json, sys, time
import torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import capi

lib = capi.load()
ctx = ctypes.c_void_p()
rc = lib.rsvdb_create(ctypes.byref(ctx), 0); assert rc == 0, rc
dev = torch.device("cuda:0")
lib.rsvdb_set_stream(ctx, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

def cm(m, n, ld=None, fill="randn"):
    """column-major m x n (leading dimension ld) as a torch tensor t of shape (n, ld); t[:, :m].T is the matrix."""
    ld = ld or m
    t = torch.empty((n, ld), dtype=torch.float64, device=dev)
    if fill == "randn": t.normal_()
    else: t.fill_(float("nan"))
    return t

def check(m, n, l, lda=None):
    A = cm(m, n, lda); X = cm(n, l); Y = cm(m, l, fill="nan")
    rc = lib.rsvdb_gemm_an_dev(ctx, A.data_ptr(), m, n, A.shape[1], X.data_ptr(), n, l, Y.data_ptr(), m)
    assert rc == 0, (rc, lib.rsvdb_last_error(ctx))
    Am = A[:, :m].T; ref = Am @ X.T
    e1 = ((Y.T - ref).norm() / ref.norm()).item() if ref.numel() else 0.0
    Q = cm(m, l); Z = cm(n, l, fill="nan"); B = cm(l, n, fill="nan")
    rc = lib.rsvdb_gemm_at_dev(ctx, A.data_ptr(), m, n, A.shape[1], Q.data_ptr(), m, l, Z.data_ptr(), n, 0); assert rc == 0, rc
    rc = lib.rsvdb_gemm_at_dev(ctx, A.data_ptr(), m, n, A.shape[1], Q.data_ptr(), m, l, B.data_ptr(), l, 1); assert rc == 0, rc
    ref2 = Am.T @ Q.T
    e2 = ((Z.T - ref2).norm() / ref2.norm()).item() if ref2.numel() else 0.0
    e3 = ((B.T - ref2.T).norm() / ref2.norm()).item() if ref2.numel() else 0.0
    ok = max(e1, e2, e3) < 1e-13
    print(json.dumps({"check": [m, n, l, lda], "err_an": e1, "err_at": e2, "err_at_T": e3, "ok": ok}), flush=True)
    return ok

allok = True
for (m, n, l, lda) in [(128, 16, 8, None), (256, 64, 16, None), (1000, 300, 20, None), (4096, 4096, 50, None), (5000, 777, 100, 5002),
                       (333, 129, 7, 334), (100, 100, 16, None), (20000, 1000, 104, None), (50000, 2000, 64, None), (130, 50, 128, None),
                       (2048, 513, 130, None), (999, 64, 33, 999), (17, 5, 3, 18)]:
    allok &= check(m, n, l, lda)
print(json.dumps({"all_ok": bool(allok)}), flush=True)

def timeit(fn, reps=5):
    rc = fn(); torch.cuda.synchronize()
    assert not isinstance(rc, int) or rc == 0, rc
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for (m, n, l) in [(200000, 20000, 100), (25000, 20000, 100), (100000, 1000, 20), (50000, 2000, 64), (4096, 4096, 50), (200000, 20000, 16)]:
    A = cm(m, n); X = cm(n, l); Y = cm(m, l); Q = cm(m, l); Z = cm(n, l)
    l0 = lib.rsvdb_launch_count(ctx)
    t_an = timeit(lambda: lib.rsvdb_gemm_an_dev(ctx, A.data_ptr(), m, n, m, X.data_ptr(), n, l, Y.data_ptr(), m))
    t_at = timeit(lambda: lib.rsvdb_gemm_at_dev(ctx, A.data_ptr(), m, n, m, Q.data_ptr(), m, l, Z.data_ptr(), n, 0))
    fl = 2.0 * m * n * l
    t_cb = timeit(lambda: torch.matmul(X, A, out=Y)) if m * n * 8 < 40e9 else float("nan")   # Y^T = X^T A^T : cuBLAS on the same data
    print(json.dumps({"time": [m, n, l], "an_ms": round(t_an, 3), "an_tflops": round(fl / t_an * 1e-9, 2), "at_ms": round(t_at, 3),
                      "at_tflops": round(fl / t_at * 1e-9, 2), "cublas_an_ms": round(t_cb, 3), "cublas_an_tflops": round(fl / t_cb * 1e-9, 2),
                      "an_GBps": round(m * n * 8 / t_an * 1e-6, 1), "at_GBps": round(m * n * 8 / t_at * 1e-6, 1)}), flush=True)
    del A, X, Y, Q, Z
lib.rsvdb_destroy(ctx)
