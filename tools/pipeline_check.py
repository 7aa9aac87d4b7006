"""GPU trip 2: end-to-end parity of the CUDA path against the oracle + first timings."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
from oracle import rsvd_oracle as O

E = Engine(0)
rng = np.random.default_rng(0)
ok_all = True
def report(name, ok, **kw):
    global ok_all
    ok_all &= bool(ok)
    print(json.dumps({"test": name, "ok": bool(ok), **{k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kw.items()}}), flush=True)

# --- QR API (TSQR under it) ---
for (m, n) in [(100, 16), (1000, 20), (5000, 50), (20000, 100), (777, 33), (300, 110), (64, 64), (250, 250), (400, 130)]:
    A = rng.standard_normal((m, n))
    if (m, n) == (1000, 20): A[:, 5:] = A[:, :5] @ rng.standard_normal((5, 15))   # rank 5
    Q, R = E.qr_decomposition_reduced(A)
    orth = np.linalg.norm(Q.T @ Q - np.eye(n)); rec = np.linalg.norm(Q @ R - A) / np.linalg.norm(A)
    tri = np.linalg.norm(np.tril(R, -1)); dpos = bool(np.all(np.diag(R) >= 0))
    report(f"qr_reduced_{m}x{n}", orth < 1e-12 and rec < 1e-13 and tri == 0 and dpos, orth=orth, rec=rec)
A = rng.standard_normal((60, 12)); Q, R = E.qr_decomposition_full(A)
report("qr_full_60x12", np.linalg.norm(Q.T @ Q - np.eye(60)) < 1e-12 and np.linalg.norm(Q @ R - A) < 1e-12, orth=np.linalg.norm(Q.T @ Q - np.eye(60)), rec=np.linalg.norm(Q @ R - A))

# --- small SVD back-ends ---
for shape in [(16, 100), (64, 64), (100, 30), (50, 300), (100, 20000), (120, 120), (7, 3)]:
    B = rng.standard_normal(shape)
    U, S, V = E.svd(B, SVDMethod.Jacobi)
    Sref = np.linalg.svd(B, compute_uv=False)
    k = min(shape)
    rel = np.max(np.abs(S - Sref) / Sref[0]); rec = np.linalg.norm(B - (U * S) @ V.T) / np.linalg.norm(B)
    orth = max(np.linalg.norm(U.T @ U - np.eye(k)), np.linalg.norm(V.T @ V - np.eye(k)))
    report(f"svd_jacobi_{shape[0]}x{shape[1]}", rel < 1e-13 and rec < 1e-13 and orth < 1e-12 and U.shape == (shape[0], k) and V.shape == (shape[1], k),
           rel=rel, rec=rec, orth=orth, sweeps=E.last_svd_info()[0])

# --- intermediate_step / rSVD vs oracle ---
def parity(name, A, l, q=2, methods=(SVDMethod.Jacobi, SVDMethod.ParallelJacobi)):
    m, n = A.shape
    Om = W.omega(n, l)
    Qo = O.intermediate_step(A, Om, l, q); Qg = E.intermediate_step(A, Om, l, q)
    orth = np.linalg.norm(Qg.T @ Qg - np.eye(l))
    for meth in methods:
        Uo, So, Vo = O.rsvd(A, Om, l, q, int(meth))
        Ug, Sg, Vg = E.rSVD(A, l, meth, Omega=Om, q=q)
        okS, relS = O.sigma_close(Sg, So)
        eo = O.reconstruction_error(A, Uo, So, Vo); eg = O.reconstruction_error(A, Ug, Sg, Vg)
        nA = np.linalg.norm(A)
        r = int(np.sum(So >= 1e-6 * So[0]))
        sin = O.subspace_sin_theta(Uo[:, :r], Ug[:, :r]) if (r == l or So[r] < 1e-3 * So[r - 1]) else float("nan")
        orthU = np.linalg.norm(Ug.T @ Ug - np.eye(Ug.shape[1])); orthV = np.linalg.norm(Vg.T @ Vg - np.eye(Vg.shape[1]))
        ok = okS and eg <= eo + 1e-8 * nA and orthU < 1e-10 and orthV < 1e-10 and orth < 1e-10 and not (sin > 1e-6)
        report(f"rsvd_{name}_{meth.name}", ok, relS=relS, err_gpu=eg, err_oracle=eo, sin_theta=sin, rank=r, orthQ=orth, orthU=orthU, orthV=orthV)

for nm, gen in W.C1_CASES:
    parity("c1_" + nm, gen(), W.C1_L)
A = rng.standard_normal((300, 120)) @ np.diag(0.8 ** np.arange(120)) @ rng.standard_normal((120, 120))
parity("decay_300x120_l20", A, 20)
parity("gauss_250x250_l100", rng.uniform(-1, 1, (250, 250)), 100)
parity("gauss_2000x500_l64_q0", rng.standard_normal((2000, 500)), 64, q=0)
parity("pod_5000x400_l64", W.c4_pod(5000, 400), 64)
parity("pca_20000x300_l20", W.c3_pca(20000, 300), 20)

# --- power back-end, PM, manualMatrixMultiply ---
B = rng.standard_normal((16, 100)) * (0.5 ** np.arange(16))[:, None]
U, S, V = E.svd(B, SVDMethod.Power)
Sref = np.linalg.svd(B, compute_uv=False)
report("svd_power_16x100", U.shape == (16, 16) and V.shape == (100, 100) and np.max(np.abs(S - Sref) / Sref) < 1e-6, rel=np.max(np.abs(S - Sref) / Sref),
       rec=np.linalg.norm(B - (U * S) @ V[:16, :]) / np.linalg.norm(B))
sg, u, v = E.PM(B)
report("pm_16x100", abs(sg - Sref[0]) / Sref[0] < 1e-8 and abs(abs(u @ B @ v) - Sref[0]) / Sref[0] < 1e-8, sigma=sg, ref=Sref[0])
A1 = rng.standard_normal((37, 53)); B1 = rng.standard_normal((53, 29))
C1 = E.manualMatrixMultiply(A1, B1)
report("manual_matmul", np.linalg.norm(C1 - A1 @ B1) / np.linalg.norm(A1 @ B1) < 1e-14)
try:
    E.manualMatrixMultiply(A1, rng.standard_normal((52, 3))); report("manual_matmul_mismatch_raises", False)
except ValueError:
    report("manual_matmul_mismatch_raises", True)
try:
    E.rSVD(A1, 5, 7); report("rsvd_bad_method_raises", False)
except ValueError:
    report("rsvd_bad_method_raises", True)
Ur, Sr, Vr = E.rSVD(rng.standard_normal((200, 80)) * 1.0, 10, SVDMethod.Power, Omega=W.omega(80, 10))
report("rsvd_power_shapes", Ur.shape == (200, 10) and Sr.shape == (10,) and Vr.shape == (80, 80))
print(json.dumps({"all_ok": bool(ok_all)}), flush=True)

# --- timings through the host API with phase breakdown ---
E.set_profiling(True)
for nm, A, l in [("c2_4096x4096_l50", W.c2_image(), 50), ("c3_100000x1000_l20", W.c3_pca(), 20), ("c4_50000x2000_l64", W.c4_pod(), 64)]:
    m, n = A.shape; Om = W.omega(n, l)
    E.rSVD(A, l, SVDMethod.Jacobi, Omega=Om); E.phase_ms()
    t0 = time.time(); U, S, V = E.rSVD(A, l, SVDMethod.Jacobi, Omega=Om); dt = time.time() - t0
    ph = E.phase_ms()
    t1 = time.time(); Uo, So, Vo = O.rsvd(A, Om, l, 2, 0); dto = time.time() - t1
    okS, relS = O.sigma_close(S, So)
    print(json.dumps({"time": nm, "e2e_ms": round(dt * 1e3, 2), "oracle_cpu_ms": round(dto * 1e3, 1), "phases_ms": {k: round(v, 3) for k, v in ph.items()},
                      "sigma_ok": okS, "relS": relS, "err_gpu": O.reconstruction_error(A, U, S, V), "err_oracle": O.reconstruction_error(A, Uo, So, Vo),
                      "sweeps": E.last_svd_info()}), flush=True)
