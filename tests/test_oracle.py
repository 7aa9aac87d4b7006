"""CPU tests: the oracle (oracle/) against (1) the golden vectors produced by the reference's own first-party sources
(tests/golden/ref_outputs.npz, see make_golden.py), (2) those sources live when oracle/_ref is present, and (3) the
mathematical known answers of BASELINE.md section 3."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import make_golden as G  # noqa: E402
from rsvd_kamaneh_raganato_terrana_b200 import workloads as W  # noqa: E402

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "ref_outputs.npz")


@pytest.mark.parametrize("name", list(G.rsvd_inputs().keys()))
def test_rsvd_matches_reference_golden(oracle, name):
    A, l = G.rsvd_inputs()[name]
    Om = W.omega(A.shape[1], l)
    Q = oracle.intermediate_step(A, Om, l, 2)
    Qref = GOLD[f"istep/{name}/Q"]
    assert Q.shape == Qref.shape
    assert np.linalg.norm(Q.T @ Q - np.eye(l)) < 1e-12
    for meth, tag in ((oracle.JACOBI, "jacobi"), (oracle.PARALLEL_JACOBI, "pjacobi")):
        U, S, V = oracle.rsvd(A, Om, l, 2, meth)
        Sref = GOLD[f"rsvd/{name}/{tag}/S"]
        # oracle (LAPACK-blocked QR, OpenBLAS products) vs reference sources (unblocked QR, plain loops): rounding-level
        # differences only; ParallelJacobi amplifies them (its stop test is loose)
        tol = 1e-10 if tag == "jacobi" else 1e-8
        assert np.all(np.abs(S - Sref) <= tol * np.maximum(Sref, 1e-6 * Sref[0])), (name, tag)
        assert tuple(V.shape) == tuple(GOLD[f"rsvd/{name}/{tag}/Vshape"])
        err, eref = oracle.reconstruction_error(A, U, S, V), float(GOLD[f"rsvd/{name}/{tag}/err"])
        assert abs(err - eref) <= 1e-9 * np.linalg.norm(A) + 1e-6 * eref
        r = int(np.sum(Sref >= 1e-6 * Sref[0]))
        if r == l or Sref[r] < 1e-3 * Sref[r - 1]:
            assert oracle.subspace_sin_theta(GOLD[f"rsvd/{name}/{tag}/U"][:, :r], U[:, :r]) < 1e-6


@pytest.mark.parametrize("name", list(G.small_inputs().keys()))
def test_small_svd_and_qr_match_reference_golden(oracle, name):
    B = G.small_inputs()[name]
    for fn, tag in ((oracle.svd_jacobi, "jacobi"), (oracle.svd_parallel_jacobi, "pjacobi")):
        U, S, V, _ = fn(B)
        np.testing.assert_allclose(S, GOLD[f"svd/{name}/{tag}/S"], rtol=0, atol=1e-13 * max(1.0, GOLD[f"svd/{name}/{tag}/S"][0]))
        assert list(U.shape) + list(V.shape) == list(GOLD[f"svd/{name}/{tag}/shapes"])
    if B.shape[0] >= B.shape[1]:
        Q, R = oracle.givens_qr(B, reduced=True)
        np.testing.assert_allclose(np.abs(R), GOLD[f"qr/{name}/absR"], atol=1e-12 * np.abs(R).max())
        np.testing.assert_allclose(np.diag(R), GOLD[f"qr/{name}/diagR"], atol=1e-12 * np.abs(R).max())


def test_scalar_helpers_match_reference_golden(oracle):
    lib = oracle._lib()
    for row, out in zip(GOLD["make_jacobi/in"], GOLD["make_jacobi/out"]):
        c = ctypes.c_double(); s = ctypes.c_double()
        ok = lib.oc_make_jacobi(ctypes.c_double(row[0]), ctypes.c_double(row[1]), ctypes.c_double(row[2]), ctypes.byref(c), ctypes.byref(s))
        assert (float(ok), c.value, s.value) == tuple(out)
    for M, out in zip(GOLD["svd2x2/in"], GOLD["svd2x2/out"]):
        o = [ctypes.c_double() for _ in range(4)]
        lib.oc_real_2x2_jacobi_svd(ctypes.c_double(M[0, 0]), ctypes.c_double(M[0, 1]), ctypes.c_double(M[1, 0]), ctypes.c_double(M[1, 1]),
                                   ctypes.c_double(np.finfo(float).tiny), *[ctypes.byref(x) for x in o])
        np.testing.assert_allclose([x.value for x in o], out, rtol=0, atol=1e-15)
    for n, s in GOLD["pm_iterations"]:
        assert oracle.pm_iterations(int(n)) == int(s)
    assert [oracle.pm_iterations(n) for n in (100, 1000, 2000, 4096, 20000)] == [148, 149, 150, 150, 151]


def test_known_answers(oracle):
    """identity inputs give sigma = 1 and ||A - U S V^T|| = sqrt(n - 16) for any Omega; the ramp
    matrix has sigma_1 = 5.77391767e5, sigma_2 = 1.44312761e3 and numerical rank 2."""
    for n in (100, 110, 140, 160):
        A = W.c1_identity(n)
        for seed in (0, 1):
            U, S, V = oracle.rsvd(A, W.omega(n, 16, seed), 16, 2, oracle.JACOBI)
            assert np.max(np.abs(S - 1.0)) < 1e-14
            assert abs(oracle.reconstruction_error(A, U, S, V) - np.sqrt(n - 16)) < 1e-12
    A = W.c1_ramp(100)
    assert abs(np.linalg.norm(A) - 577393.5702794065) < 1e-6
    for seed in (0, 1, 2):
        U, S, V = oracle.rsvd(A, W.omega(100, 16, seed), 16, 2, oracle.JACOBI)
        assert abs(S[0] - 5.77391767e5) / 5.77391767e5 < 1e-8 and abs(S[1] - 1.44312761e3) / 1.44312761e3 < 1e-8
        assert S[2] < 1e-9 and oracle.reconstruction_error(A, U, S, V) < 1e-7


def test_power_backend_shapes_and_accuracy(oracle):
    rng = np.random.default_rng(5)
    B = rng.standard_normal((16, 100)) * (0.5 ** np.arange(16))[:, None]
    U, S, V, info = oracle.svd_power(B, 0, seed=1)
    assert U.shape == (16, 16) and V.shape == (100, 100) and info["found"] == 16      # V: n x n, vectors in rows
    Sref = np.linalg.svd(B, compute_uv=False)
    assert np.max(np.abs(S - Sref) / Sref) < 1e-9
    assert np.linalg.norm(B - (U * S) @ V[:16, :]) / np.linalg.norm(B) < 1e-9


def test_unsupported_method_raises(oracle):
    with pytest.raises(ValueError):
        oracle.rsvd(np.eye(10), W.omega(10, 4), 4, 2, 7)
    with pytest.raises(ValueError):
        oracle.manual_matmul(np.ones((3, 4)), np.ones((5, 2)))


def test_oracle_vs_reference_sources_live(oracle):
    """Only where oracle/_ref exists (dev container, or a box that received the prebuilt .so)."""
    if not oracle.RefLib.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    ref = oracle.RefLib()
    rng = np.random.default_rng(11)
    for shape in [(12, 40), (33, 33), (45, 9)]:
        B = rng.standard_normal(shape)
        for fo, meth in ((oracle.svd_jacobi, 0), (oracle.svd_parallel_jacobi, 2)):
            U, S, V, _ = fo(B); U2, S2, V2 = ref.svd(B, meth)
            assert np.array_equal(S, S2) and np.array_equal(U, U2) and np.array_equal(V, V2)   # bit-identical restatement
    A = rng.standard_normal((90, 40)); Om = rng.standard_normal((40, 8))
    assert oracle.subspace_sin_theta(oracle.intermediate_step(A, Om, 8, 2), ref.intermediate_step(A, Om, 8, 2)) < 1e-12
    Q1, R1 = oracle.givens_qr(A[:30, :7]); Q2, R2 = ref.qr_reduced(A[:30, :7])
    assert np.array_equal(Q1, Q2) and np.array_equal(R1, R2)
    np.testing.assert_allclose(oracle.manual_matmul(A, Om), ref.manual_matmul(A, Om), rtol=0, atol=0)
    with pytest.raises(ValueError):
        ref.manual_matmul(A, A)
    with pytest.raises(ValueError):
        ref.rsvd(A, Om, 8, 5)


# ---------------------------------------------------------------------------------------------------------------------
# PCA front / back steps: the numpy restatement vs the reference's own PCA class (golden)
# ---------------------------------------------------------------------------------------------------------------------
def pca_cases():
    g = GOLD
    d = G.pca_inputs()
    for nm in ("tourists", "athletic"):          # reference datasets travel inside the fixture
        d[nm] = np.asfortranarray(g[f"pca/{nm}/data"])
    return d, g


@pytest.mark.parametrize("name", ["tourists", "athletic", "offset_500x60", "wide_30x50"])
@pytest.mark.parametrize("normalize", [0, 1])
def test_pca_restatement_matches_reference_class(oracle, name, normalize):
    d, g = pca_cases()
    D = d[name]
    for meth, tag in ((oracle.JACOBI, "jacobi"), (oracle.PARALLEL_JACOBI, "pjacobi")):
        p = oracle.PCA(D, bool(normalize), meth)
        key = f"pca/{name}/n{normalize}/{tag}/"
        ev = g[key + "explained_variance"]
        tol = 1e-6 if tag == "pjacobi" else 1e-10                      # ParallelJacobi stops early
        assert np.max(np.abs(p.explainedVariance() - ev)) <= tol * ev[0]
        assert np.max(np.abs(p.explainedVarianceRatio() - g[key + "ratio"])) <= tol
        np.testing.assert_allclose(p.mean, g[key + "mean"], rtol=1e-13, atol=1e-13 * np.abs(D).max())
        if ev[-1] > 1e-8 * ev[0]:     # a numerically zero component has an arbitrary direction (centred 30 x 50 data has rank 29)
            rec = p.reconstructFromPCA(p.projectToPCA(D[:7]))
            np.testing.assert_allclose(rec, g[key + "reconstruct"], rtol=0, atol=(1e-5 if tag == "pjacobi" else 1e-9) * np.abs(D).max())
        # scores / projections agree up to the sign of each well-separated component
        sep = np.r_[np.abs(np.diff(ev)) > 1e-6 * ev[0], True] & np.r_[True, np.abs(np.diff(ev)) > 1e-6 * ev[0]] & (ev > 1e-8 * ev[0])
        sc = np.abs(p.scores())[:, sep]; ref_sc = g[key + "abs_scores"][:, sep]
        assert np.max(np.abs(sc - ref_sc)) <= 1e-5 * ev[0] * np.sqrt(D.shape[0])


def test_pca_rejects_degenerate_input(oracle):
    with pytest.raises(ValueError, match="at least 2 rows and 2 columns"):
        oracle.PCA(np.zeros((1, 5)))
    if oracle.RefLib.available():
        with pytest.raises(ValueError, match="at least 2 rows and 2 columns"):
            oracle.RefLib().pca(np.zeros((5, 1)))


# ---------------------------------------------------------------------------------------------------------------------
# POD wrappers: the numpy restatement vs the reference's own POD class (golden)
# ---------------------------------------------------------------------------------------------------------------------
def _pod_compare(Wm, sg, absW_ref, sg_ref, r, loose):
    tol = 1e-6 if loose else 1e-9
    assert sg.shape == sg_ref.shape and Wm.shape == absW_ref.shape          # same N from the energy criterion
    assert np.max(np.abs(sg[:r] - sg_ref[:r])) <= tol * sg_ref[0]
    # modes agree up to sign where the singular values are separated and not at noise level
    s = sg_ref[:Wm.shape[1]]
    gap = np.r_[np.abs(np.diff(sg_ref[:Wm.shape[1] + 1]))[: len(s)], np.inf][: len(s)] if len(sg_ref) > len(s) else np.r_[np.abs(np.diff(s)), np.inf]
    ok = (gap > 1e-4 * sg_ref[0]) & (s > 1e-9 * sg_ref[0])
    ok[1:] &= np.abs(np.diff(s)) > 1e-4 * sg_ref[0]
    if ok.any():
        scale = np.abs(absW_ref[:, ok]).max()
        assert np.max(np.abs(np.abs(Wm[:, ok]) - absW_ref[:, ok])) <= (1e-3 if loose else 1e-6) * scale


@pytest.mark.parametrize("name", list(G.pod_inputs().keys()))
def test_pod_restatement_matches_reference_class(oracle, name):
    S, Xh, D, r, tol = G.pod_inputs()[name]
    for variant, st in G.POD_CASES:
        key = f"pod/{name}/v{variant}/t{st}/"
        if key + "sigma" not in GOLD.files:
            continue
        Wm, sg = oracle.pod(variant, S, r, tol, st, Xh if variant >= 2 else None, D if variant == 3 else None,
                            G.pod_omega(S, variant, r) if st >= 3 else None)
        _pod_compare(Wm, sg, GOLD[key + "absW"], GOLD[key + "sigma"], r, loose=st in (2, 5))


def test_pod_bad_svd_type(oracle):
    with pytest.raises(ValueError, match=r"svd_type should be in \[0,5\]"):
        oracle.pod(1, np.eye(6), 2, 1e-3, 7)
