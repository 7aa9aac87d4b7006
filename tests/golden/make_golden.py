"""Generates tests/golden/ref_outputs.npz from the reference's OWN first-party sources, compiled over the Eigen/MPI
stand-in by oracle/Makefile (oracle/_ref/libref_rsvd.so).  Run in the dev container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures are what the -m "not gpu" tests pin the oracle against, and what the -m gpu tests compare the CUDA path
with, on boxes where /root/reference is absent.  Inputs are regenerated from seeds (workloads.py), only outputs are stored.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import rsvd_oracle as O  # noqa: E402
from rsvd_kamaneh_raganato_terrana_b200 import workloads as W  # noqa: E402


def small_inputs():
    """Seeded small matrices for the SVD / QR back-ends (name -> matrix)."""
    rng = np.random.default_rng(20260101)
    d = {
        "gauss_16x100": rng.standard_normal((16, 100)),
        "gauss_64x64": rng.standard_normal((64, 64)),
        "gauss_100x30": rng.standard_normal((100, 30)),
        "gauss_50x300": rng.standard_normal((50, 300)),
        "decay_40x40": rng.standard_normal((40, 40)) @ np.diag(0.5 ** np.arange(40)) @ rng.standard_normal((40, 40)),
        "diag_1_to_100": np.diag(np.arange(1.0, 101.0)),                      # image_compression/data/input/mat: sigma = sorted |diag|
        "qr_test2_4x3": np.arange(1.0, 13.0).reshape(4, 3),                  # image_compression/tests/QR_test2.cpp:24-29
    }
    return {k: np.asfortranarray(v) for k, v in d.items()}


def rsvd_inputs():
    rng = np.random.default_rng(20260102)
    d = {nm: (gen(), W.C1_L) for nm, gen in W.C1_CASES}                      # tests/rSVD_test.cpp on input/*.mtx, l = 16
    d["decay_300x120"] = (np.asfortranarray(rng.standard_normal((300, 120)) @ np.diag(0.8 ** np.arange(120)) @ rng.standard_normal((120, 120))), 20)
    d["uniform_250x250"] = (np.asfortranarray(rng.uniform(-1, 1, (250, 250))), 50)   # tests/rSVD_test2.cpp shape, l = 50
    d["pod_2000x300"] = (W.c4_pod(2000, 300), 32)
    return d


def _parse_table(path, skip_cols):
    """The loaders of PCA/tests/pca_test.cpp:7-57 (tourists: header + 3 label columns) and athletic_test.cpp (header + 1)."""
    rows = []
    with open(path) as f:
        next(f)
        for line in f:
            vals = []
            for tok in line.split()[skip_cols:]:
                try:
                    vals.append(float(tok))
                except ValueError:
                    pass
            if vals:
                rows.append(vals)
    w = max(len(r) for r in rows)
    return np.asfortranarray(np.array([r for r in rows if len(r) == w]))


def pca_inputs():
    """name -> (data, rows to project).  The two reference datasets are stored in the fixture (the GPU box has no
    /root/reference); the synthetic ones are regenerated from seeds."""
    rng = np.random.default_rng(20260103)
    ref_data = Path("/root/reference/PCA/data/input")
    d = {}
    if ref_data.is_dir():
        d["tourists"] = _parse_table(ref_data / "tourists.txt", 3)             # PCA/tests/pca_test.cpp
        d["athletic"] = _parse_table(ref_data / "dataset_athletic.txt", 1)     # PCA/tests/athletic_test.cpp
    d["offset_500x60"] = np.asfortranarray(W.c3_pca(500, 60, seed=31) * (1.0 + np.arange(60) % 5) + 3.0 * rng.standard_normal(60))
    d["wide_30x50"] = np.asfortranarray(rng.standard_normal((30, 50)) + rng.uniform(-2, 2, 50))
    return d


def pod_inputs():
    """name -> (S, Xh, D, r, tol).  Xh: SPD tridiagonal 'stiffness + mass'-like operator, D: positive quadrature weights."""
    rng = np.random.default_rng(20260104)
    def spd(n):
        return np.asfortranarray(2.0 * np.eye(n) - 0.5 * np.eye(n, k=1) - 0.5 * np.eye(n, k=-1))
    def weights(n):
        return np.asfortranarray(np.diag(1.0 + 0.5 * (np.arange(n) % 3)))
    tall = np.asfortranarray(rng.standard_normal((300, 30)) @ np.diag(0.6 ** np.arange(30)) @ rng.standard_normal((30, 40)))
    d = {
        "heat_400x60": (W.c4_pod(400, 60), spd(400), weights(60), 12, 1e-6),
        "decay_300x40": (tall, spd(300), weights(40), 10, 1e-2),
        "wide_40x300": (np.asfortranarray(tall.T), spd(40), weights(300), 10, 1e-2),       # ns > Nh branches
    }
    return d


POD_CASES = [(v, t) for v in (0, 1, 2, 3) for t in (1, 2, 4, 5)]


def pod_omega(S, variant, r):
    Nh, ns = S.shape
    return W.omega(ns if variant == 0 else min(Nh, ns), r)


def main():
    ref = O.RefLib()
    out = {}
    for nm, (S, Xh, D, r, tol) in pod_inputs().items():
        for variant, st in POD_CASES:
            if S.shape[1] > S.shape[0] and variant >= 2 and st <= 2:
                continue      # the reference writes past U there (POD.cpp:300-302 loops over Utilde.cols() = Nh > r)
            Wm, sg = ref.pod(variant, S, r, tol, st, Xh if variant >= 2 else None, D if variant == 3 else None,
                             pod_omega(S, variant, r) if st >= 3 else None)
            out[f"pod/{nm}/v{variant}/t{st}/sigma"] = sg
            out[f"pod/{nm}/v{variant}/t{st}/absW"] = np.abs(Wm)
    for nm, D in pca_inputs().items():
        if nm in ("tourists", "athletic"):
            out[f"pca/{nm}/data"] = D
        for norm in (0, 1):
            for meth, tag in ((O.JACOBI, "jacobi"), (O.PARALLEL_JACOBI, "pjacobi")):
                r = ref.pca(D, bool(norm), meth, project=D[:7])
                for key in ("explained_variance", "ratio", "mean"):
                    out[f"pca/{nm}/n{norm}/{tag}/{key}"] = r[key]
                out[f"pca/{nm}/n{norm}/{tag}/abs_scores"] = np.abs(r["scores"])
                out[f"pca/{nm}/n{norm}/{tag}/abs_project"] = np.abs(r["project"])
                out[f"pca/{nm}/n{norm}/{tag}/reconstruct"] = r["reconstruct"]
    for nm, (A, l) in rsvd_inputs().items():
        Om = W.omega(A.shape[1], l)
        Q = ref.intermediate_step(A, Om, l, 2)
        out[f"istep/{nm}/Q"] = Q
        for meth, tag in ((O.JACOBI, "jacobi"), (O.PARALLEL_JACOBI, "pjacobi")):
            U, S, V = ref.rsvd(A, Om, l, meth)
            out[f"rsvd/{nm}/{tag}/S"] = S
            out[f"rsvd/{nm}/{tag}/err"] = np.array(O.reconstruction_error(A, U, S, V))
            out[f"rsvd/{nm}/{tag}/U"] = U
            out[f"rsvd/{nm}/{tag}/Vshape"] = np.array(V.shape)
    for nm, B in small_inputs().items():
        for meth, tag in ((O.JACOBI, "jacobi"), (O.PARALLEL_JACOBI, "pjacobi")):
            U, S, V = ref.svd(B, meth)
            out[f"svd/{nm}/{tag}/S"] = S
            out[f"svd/{nm}/{tag}/shapes"] = np.array(list(U.shape) + list(V.shape))
        if B.shape[0] >= B.shape[1]:
            Q, R = ref.qr_reduced(B)
            out[f"qr/{nm}/absR"] = np.abs(R)
            out[f"qr/{nm}/diagR"] = np.diag(R).copy()
    # scalar helpers
    rng = np.random.default_rng(7)
    xyz = rng.standard_normal((20, 3)); xyz[3] = [1.0, 0.0, 2.0]
    out["make_jacobi/in"] = xyz
    out["make_jacobi/out"] = np.array([[float(ok), c, s] for ok, c, s in (ref.make_jacobi(*row) for row in xyz)])
    M = rng.standard_normal((10, 2, 2))
    out["svd2x2/in"] = M
    out["svd2x2/out"] = np.array([ref.real_2x2_jacobi_svd(m) for m in M])
    out["pm_iterations"] = np.array([[n, O.pm_iterations(n)] for n in (100, 1000, 2000, 4096, 20000)])
    np.savez_compressed(Path(__file__).with_name("ref_outputs.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
