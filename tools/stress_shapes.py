"""Randomised shape sweep of the host rSVD entry point against the CPU oracle (sigma parity, reconstruction, orthogonality)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
from oracle import rsvd_oracle as O
E = Engine(0)
rng = np.random.default_rng(2026)
cases = [(50, 3000, 8, 2), (3000, 50, 50, 2), (777, 1234, 1, 2), (2048, 2048, 128, 1), (5000, 700, 150, 2), (1500, 1500, 200, 0), (129, 257, 100, 3),
         (10000, 300, 104, 2), (640, 640, 101, 2), (4100, 90, 90, 2), (33, 33, 33, 2), (2, 2, 1, 2), (9000, 2000, 96, 2), (600, 5000, 112, 1)]
bad = 0
for (m, n, l, q) in cases:
    r = min(m, n, 60)
    A = np.asfortranarray(rng.standard_normal((m, r)) @ np.diag(0.75 ** np.arange(r)) @ rng.standard_normal((r, n)) + 1e-7 * rng.standard_normal((m, n)))
    Om = W.omega(n, l)
    try:
        U, S, V = E.rSVD(A, l, SVDMethod.Jacobi, Omega=Om, q=q)
    except Exception as e:
        print(json.dumps({"case": [m, n, l, q], "error": str(e)[:200]})); bad += 1; continue
    Uo, So, Vo = O.rsvd(A, Om, l, q, O.JACOBI)
    k = min(l, n)
    okS, rel = O.sigma_close(S, So)
    eg, eo = O.reconstruction_error(A, U, S, V), O.reconstruction_error(A, Uo, So, Vo)
    kk = min(k, m)
    orthU = float(np.linalg.norm(U[:, :kk].T @ U[:, :kk] - np.eye(kk))); orthV = float(np.linalg.norm(V.T @ V - np.eye(V.shape[1])))
    ok = bool(okS and abs(eg - eo) <= 1e-8 * np.linalg.norm(A) and (orthU < 1e-9 or m < k) and orthV < 1e-9 and U.shape == Uo.shape and V.shape == Vo.shape)
    bad += 0 if ok else 1
    print(json.dumps({"case": [m, n, l, q], "ok": ok, "relS": rel, "err_gpu": eg, "err_oracle": eo, "orthU": orthU, "orthV": orthV, "shapes": [list(U.shape), list(Uo.shape)]}), flush=True)
print(json.dumps({"failures": bad}))
