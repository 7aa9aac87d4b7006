"""CPU tests of the host-side logic: row sharding, deterministic workloads, MatrixMarket IO and -- with two gloo ranks --
the communication pattern of the row-sharded rSVD (all-gather of TSQR R factors, all-reduce of A^T Q partial sums)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rsvd_kamaneh_raganato_terrana_b200 import mtx, workloads as W  # noqa: E402


def test_row_split_matches_reference_rule():
    # rows/P (+1 for the first rows%P ranks), reference src/rSVD.cpp:20-23
    for m, P in [(200000, 8), (10, 3), (7, 7), (5, 8), (100, 1)]:
        parts = [W.row_split(m, P, r) for r in range(P)]
        assert parts[0][0] == 0 and sum(rows for _, rows in parts) == m
        for (o1, r1), (o2, _) in zip(parts, parts[1:]):
            assert o1 + r1 == o2
        assert max(r for _, r in parts) - min(r for _, r in parts) <= 1


def test_workloads_are_deterministic_and_shaped():
    assert np.array_equal(W.omega(50, 7), W.omega(50, 7))
    A = W.c1_ramp(100)
    assert A[0, 0] == 1.0 and A[99, 99] == 100 * 99 + 100 and np.linalg.matrix_rank(A) == 2
    assert W.c3_pca(500, 60).shape == (500, 60) and abs(W.c3_pca(500, 60).mean(axis=0)).max() < 1e-12
    P = W.c4_pod(400, 50)
    s = np.linalg.svd(P, compute_uv=False)
    assert s[30] < 1e-8 * s[0]                       # numerically rank-deficient, like real POD snapshots
    rp, ci, v = W.c4_sparse(1000, 10)
    assert rp[-1] == len(ci) == len(v) == 11000 and ci.dtype == np.int32 and rp.dtype == np.int64


def test_matrix_market_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((7, 5)); A[A < 0.3] = 0
    p = tmp_path / "a.mtx"
    mtx.save_coordinate(p, A, tol=1e-300)
    assert np.array_equal(mtx.load_dense(p), A)
    m, n, rowptr, col, val = mtx.load_csr(p)
    D = np.zeros((m, n))
    for i in range(m):
        D[i, col[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
    assert np.array_equal(D, A)
    q = tmp_path / "s.mtx"
    mtx.save_array(q, np.arange(4.0))
    assert np.array_equal(mtx.load_dense(q).ravel(), np.arange(4.0))
    # the reference's input files are "coordinate real general", 1-based, densified on load (tests/rSVD_test.cpp:54-57)
    (tmp_path / "i.mtx").write_text("%%MatrixMarket matrix coordinate real general\n3 3 3\n1 1 1.0\n2 2 1.0\n3 3 1.0\n")
    assert np.array_equal(mtx.load_dense(tmp_path / "i.mtx"), np.eye(3))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sharded_worker(rank, world, port, m, n, l, q, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import rsvd_oracle as O
    import sharded_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    A = rng.standard_normal((m, 40)) @ np.diag(0.7 ** np.arange(40)) @ rng.standard_normal((40, n))
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    U_p, S, V = sharded_model.rsvd_sharded(A[off:off + rows], Om, l, q, dist, torch, O)
    np.savez(Path(out_dir) / f"r{rank}.npz", U=U_p, S=S, V=V, off=off)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rsvd_two_gloo_ranks(tmp_path, oracle):
    """The multi-GPU path's algorithm (pipeline.cu: qr_inplace sharded branch, gemm_at_phase all-reduce) restated with
    numpy + gloo on 2 CPU ranks must reproduce the single-rank oracle."""
    import torch.multiprocessing as mp
    m, n, l, q, world = 301, 120, 12, 2, 2
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, m, n, l, q, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    U = np.vstack([p["U"] for p in parts]); S = parts[0]["S"]; V = parts[0]["V"]
    assert np.array_equal(parts[0]["S"], parts[1]["S"]) and np.array_equal(parts[0]["V"], parts[1]["V"])   # replicated, bit-identical
    rng = np.random.default_rng(99)
    A = rng.standard_normal((m, 40)) @ np.diag(0.7 ** np.arange(40)) @ rng.standard_normal((40, n))
    Uo, So, Vo = oracle.rsvd(A, W.omega(n, l), l, q, oracle.JACOBI)
    assert np.all(np.abs(S - So) <= 1e-10 * So[0])
    assert np.linalg.norm(U.T @ U - np.eye(l)) < 1e-12
    assert abs(oracle.reconstruction_error(A, U, S, V) - oracle.reconstruction_error(A, Uo, So, Vo)) < 1e-9 * np.linalg.norm(A)
    assert oracle.subspace_sin_theta(Uo, U) < 1e-8


def _guarded_worker(rank, world, port, m, n, l, q, rank_def, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import rsvd_oracle as O
    import sharded_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = _guarded_case(m, n, rank_def)
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    state = {}
    U_p, S, V = sharded_model.rsvd_sharded_guarded(A[off:off + rows], Om, l, q, dist, torch, O, state)
    np.savez(Path(out_dir) / f"r{rank}.npz", U=U_p, S=S, V=V, off=off, fast=state.get("fast", 0), householder=state.get("householder", 0))
    dist.barrier()
    dist.destroy_process_group()


def _guarded_case(m, n, rank_def):
    rng = np.random.default_rng(7)
    if rank_def:
        return rng.standard_normal((m, 3)) @ rng.standard_normal((3, n))
    return rng.standard_normal((m, 40)) @ np.diag(0.8 ** np.arange(40)) @ rng.standard_normal((40, n)) + 1e-3 * rng.standard_normal((m, n))


@pytest.mark.parametrize("rank_def", [False, True])
def test_sharded_guarded_cholqr2_two_gloo_ranks(tmp_path, oracle, rank_def):
    """cholqr.cu over row shards, restated with numpy + gloo: all-reduced Gram matrices, identical guard decisions on every rank
    (a rank-deficient matrix sends BOTH ranks to the TSQR at the first sketch and keeps them there), same factorisation as the oracle."""
    import torch.multiprocessing as mp
    m, n, l, q, world = 301, 120, 12, 2, 2
    port = _free_port()
    mp.spawn(_guarded_worker, args=(world, port, m, n, l, q, rank_def, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    U = np.vstack([p["U"] for p in parts]); S = parts[0]["S"]; V = parts[0]["V"]
    assert np.array_equal(parts[0]["S"], parts[1]["S"]) and np.array_equal(parts[0]["V"], parts[1]["V"])   # replicated, bit-identical
    counts = [(int(p["fast"]), int(p["householder"])) for p in parts]
    assert counts[0] == counts[1] == ((0, 2 * q + 2) if rank_def else (2 * q + 2, 0))
    A = _guarded_case(m, n, rank_def)
    Uo, So, Vo = oracle.rsvd(A, W.omega(n, l), l, q, oracle.JACOBI)
    assert np.all(np.abs(S - So) <= 1e-10 * So[0])
    r = 3 if rank_def else l
    assert np.linalg.norm(U.T @ U - np.eye(l)) < 1e-12
    assert abs(oracle.reconstruction_error(A, U, S, V) - oracle.reconstruction_error(A, Uo, So, Vo)) < 1e-9 * np.linalg.norm(A)
    assert oracle.subspace_sin_theta(Uo[:, :r], U[:, :r]) < 1e-8


def test_chol_inv_model_matches_lapack():
    """tests/sharded_model._chol_inv restates csrc/cholqr.cu k_chol_inv (right-looking Cholesky that builds L^-1 alongside L and reports
    the first non-positive pivot and ||G - I||_F^2): check it against LAPACK so that the gloo model above rests on something."""
    sys.path.insert(0, str(ROOT / "tests"))
    import sharded_model
    rng = np.random.default_rng(3)
    for l in (1, 2, 7, 33, 100):
        Y = rng.standard_normal((4 * l + 5, l))
        G = Y.T @ Y
        bad, dev2, X, R = sharded_model._chol_inv(G)
        assert bad == 0
        assert np.allclose(dev2, np.linalg.norm(G - np.eye(l)) ** 2, rtol=1e-12)
        assert np.allclose(R, np.linalg.cholesky(G).T, rtol=1e-10, atol=1e-12) and np.all(np.tril(R, -1) == 0)
        assert np.linalg.norm(X @ R - np.eye(l)) < 1e-10 * np.linalg.cond(R) and np.all(np.tril(X, -1) == 0)
        assert np.linalg.norm(X.T @ G @ X - np.eye(l)) < 1e-9
    Y = rng.standard_normal((50, 6)); Y[:, 4] = Y[:, 1] - 2.0 * Y[:, 3]          # exactly dependent column: breakdown at or after it
    bad, _, X, R = sharded_model._chol_inv(Y.T @ Y)
    assert bad == 0 or bad >= 5
    Z = np.zeros((10, 3))
    assert sharded_model._chol_inv(Z.T @ Z)[0] == 1


def _rpca_worker(rank, world, port, m, n, l, q, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import rsvd_oracle as O
    import sharded_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = W.c3_pca(m, n, seed=9) * (1.0 + (np.arange(n) % 7)) + 4.0 * np.random.default_rng(5).standard_normal(n)
    off, rows = W.row_split(m, world, rank)
    mean, inv_sd, U_p, S, V = sharded_model.rpca_sharded(A[off:off + rows], W.omega(n, l), l, q, True, dist, torch, O)
    np.savez(Path(out_dir) / f"p{rank}.npz", U=U_p, S=S, V=V, mean=mean, inv_sd=inv_sd)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_randomized_pca_two_gloo_ranks(tmp_path, oracle):
    """The row-sharded implicit-centring path (pca.cu column_stats + pipeline.cu Centering) restated with numpy + gloo on
    2 CPU ranks equals the oracle's rSVD of the explicitly centred and scaled matrix."""
    import torch.multiprocessing as mp
    m, n, l, q, world = 401, 90, 10, 2, 2
    mp.spawn(_rpca_worker, args=(world, _free_port(), m, n, l, q, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"p{r}.npz") for r in range(world)]
    U = np.vstack([p["U"] for p in parts]); S = parts[0]["S"]; V = parts[0]["V"]
    assert np.array_equal(parts[0]["S"], parts[1]["S"]) and np.array_equal(parts[0]["mean"], parts[1]["mean"])
    A = W.c3_pca(m, n, seed=9) * (1.0 + (np.arange(n) % 7)) + 4.0 * np.random.default_rng(5).standard_normal(n)
    C = A - A.mean(axis=0); C = C / np.sqrt((C * C).sum(axis=0) / (m - 1))
    np.testing.assert_allclose(parts[0]["mean"], A.mean(axis=0), rtol=1e-13)
    Uo, So, Vo = oracle.rsvd(C, W.omega(n, l), l, q, oracle.JACOBI)
    assert oracle.sigma_close(S, So)[0]
    assert abs(oracle.reconstruction_error(C, U, S, V) - oracle.reconstruction_error(C, Uo, So, Vo)) < 1e-9 * np.linalg.norm(C)
    assert np.linalg.norm(U.T @ U - np.eye(l)) < 1e-11


def _csr_worker(rank, world, port, m, n, l, q, out_dir):
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import rsvd_oracle as O
    import sharded_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = _csr_case(m, n)
    off, rows = W.row_split(m, world, rank)
    U_p, S, V = sharded_model.rsvd_sharded(M[off:off + rows], W.omega(n, l), l, q, dist, torch, O)     # the shard stays a scipy CSR block
    np.savez(Path(out_dir) / f"c{rank}.npz", U=U_p, S=S, V=V)
    dist.barrier()
    dist.destroy_process_group()


def _csr_case(m, n):
    import scipy.sparse as sp
    M = sp.random(m, n, density=0.03, format="csr", random_state=np.random.default_rng(3), data_rvs=np.random.default_rng(4).standard_normal)
    return (M + sp.eye(m, n, format="csr")).tocsr()


def test_sharded_csr_rsvd_two_gloo_ranks(tmp_path, oracle):
    """Row-sharded CSR (spmm.cu rsvd_csr_device with c->nranks > 1: shard-local SpMM / SpMM^T, all-reduce of the A^T Q partial
    sums, distributed TSQR) restated with scipy.sparse + gloo on 2 CPU ranks equals the oracle on the densified matrix."""
    import torch.multiprocessing as mp
    m, n, l, q, world = 501, 160, 14, 2, 2
    port = _free_port()
    mp.spawn(_csr_worker, args=(world, port, m, n, l, q, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"c{r}.npz") for r in range(world)]
    U = np.vstack([p["U"] for p in parts]); S = parts[0]["S"]; V = parts[0]["V"]
    assert np.array_equal(parts[0]["S"], parts[1]["S"]) and np.array_equal(parts[0]["V"], parts[1]["V"])
    A = np.asfortranarray(_csr_case(m, n).toarray())
    Uo, So, Vo = oracle.rsvd(A, W.omega(n, l), l, q, oracle.JACOBI)
    assert np.all(np.abs(S - So) <= 1e-10 * So[0])
    assert np.linalg.norm(U.T @ U - np.eye(l)) < 1e-12
    assert abs(oracle.reconstruction_error(A, U, S, V) - oracle.reconstruction_error(A, Uo, So, Vo)) < 1e-9 * np.linalg.norm(A)
