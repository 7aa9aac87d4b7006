"""SURVEY 8(f) callers at BASELINE.json's config sizes, host-pointer entry points (one upload, result download inside the
timed region) and the CPU oracle on the same input:
  POD    : standard_POD on 50000 x 2000 snapshots, r = 64, rSVD/Jacobi back-end (svd_type 4)        -- config 4's caller
  Image  : normalize + compress(k = 40 -> l = 50) + reconstruct + deNormalize of a 4096 x 4096 picture -- config 2's caller"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, workloads as W
from oracle import rsvd_oracle as O
E = Engine(0)

def best(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), r

# ---- POD
Nh, ns, r, tol = 50000, 2000, 64, 1e-4
S = W.c4_pod(Nh, ns); Om = W.omega(ns, r)
ms, (Wg, sg) = best(lambda: E.pod(1, S, r, tol, 4, Omega=Om))
t0 = time.perf_counter(); Wo, so = O.pod(1, S, r, tol, 4, Omega=Om); cpu = (time.perf_counter() - t0) * 1e3
E.set_profiling(True); E.phase_ms(); E.pod(1, S, r, tol, 4, Omega=Om); ph = {k: round(v, 3) for k, v in E.phase_ms().items()}; E.set_profiling(False)
print(json.dumps({"case": f"standard_POD {Nh}x{ns} r={r} svd_type=4", "host_call_ms": round(ms, 2), "cpu_oracle_ms": round(cpu, 1), "speedup": round(cpu / ms, 1),
                  "N": int(Wg.shape[1]), "N_oracle": int(Wo.shape[1]), "sigma_parity": O.sigma_close(sg, so)[0], "phases_ms": ph,
                  "correlation_flops": 2.0 * Nh * ns * ns, "h2d_bytes": 8 * Nh * ns}), flush=True)

# ---- Image
A = np.asfortranarray(np.round(W.c2_image() * 255.0)); m, n = A.shape; k = 40; l = k + 10
Om = W.omega(n, l)
def gpu_image():
    U, Sv, V, lo, hi, deg = E.image_compress(A, k, True, Om)
    return E.image_reconstruct(U, Sv, V, True, lo, hi), Sv
ms, (rec, Sv) = best(gpu_image)
if "--fast" in sys.argv:           # the CPU oracle of this case (power-method SVD over a 4096^2 Gram matrix, one core) takes ~22 minutes
    print(json.dumps({"case": f"Image normalize+compress(k={k})+reconstruct+deNormalize {m}x{n}", "host_call_ms": round(ms, 2), "cpu_oracle_ms": None,
                      "rel_err_gpu": float(np.linalg.norm(A - rec) / np.linalg.norm(A))}), flush=True)
    sys.exit(0)
t0 = time.perf_counter()
An, lo, hi = O.image_normalize(A); Uo, So, Vo = O.image_compress(An, k, Om); reco = O.image_denormalize(O.image_reconstruct(Uo, So, Vo), lo, hi)
cpu = (time.perf_counter() - t0) * 1e3
print(json.dumps({"case": f"Image normalize+compress(k={k})+reconstruct+deNormalize {m}x{n}", "host_call_ms": round(ms, 2), "cpu_oracle_ms": round(cpu, 1),
                  "speedup": round(cpu / ms, 1), "rel_err_gpu": float(np.linalg.norm(A - rec) / np.linalg.norm(A)),
                  "rel_err_oracle": float(np.linalg.norm(A - reco) / np.linalg.norm(A)), "sigma_head_rel_diff": float(np.max(np.abs(Sv[:3] - So[:3]) / So[:3]))}), flush=True)
