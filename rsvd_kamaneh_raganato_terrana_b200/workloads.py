"""Synthetic inputs for the five BASELINE.json configs.  Pure data generation: numpy on the host
for the small configs, torch on the device for the ones that only ever live in HBM.  Seeds are fixed so that the oracle
and the CUDA path see bit-identical inputs."""
from __future__ import annotations

import numpy as np

OMEGA_SEED = 1234


def omega(n: int, l: int, seed: int = OMEGA_SEED) -> np.ndarray:
    """Host-supplied test matrix, identical for the oracle and the CUDA path."""
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((n, l)))


# ---- C1: the reference's own inputs (input/*.mtx), regenerated from their definition -------------------------------
def c1_ramp(n: int = 100) -> np.ndarray:
    """input/sparse_matrix.mtx: dense n x n, a_ij = 100 i + j + 1 (0-based), rank 2 (python/matrix_maker.py:16-25)."""
    i = np.arange(n)[:, None]; j = np.arange(n)[None, :]
    return np.asfortranarray((100.0 * i + j + 1.0))


def c1_identity(n: int) -> np.ndarray:
    """input/sparse_matrix{100,110,140,160}.mtx are identity matrices."""
    return np.asfortranarray(np.eye(n))


C1_CASES = [("sparse_matrix", lambda: c1_ramp(100)), ("sparse_matrix100", lambda: c1_identity(100)),
            ("sparse_matrix110", lambda: c1_identity(110)), ("sparse_matrix140", lambda: c1_identity(140)),
            ("sparse_matrix160", lambda: c1_identity(160))]
C1_L = 16


# ---- C2: image-shaped 4096 x 4096, l = 50 ---------------------------------------------------------------------------
def c2_image(n: int = 4096, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    i = np.arange(n)[:, None]; j = np.arange(n)[None, :]
    G = (rng.standard_normal((n, 32)) @ rng.standard_normal((32, n))) / np.sqrt(32.0)
    A = 0.5 + 0.25 * np.sin(2 * np.pi * i / 512.0) * np.cos(2 * np.pi * j / 384.0) + 0.15 * G + 0.02 * rng.standard_normal((n, n))
    return np.asfortranarray(np.clip(A, 0.0, 1.0))


# ---- C3: PCA-shaped 100000 x 1000, l = 20 ---------------------------------------------------------------------------
def c3_pca(m: int = 100000, n: int = 1000, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((m, 40))
    W, _ = np.linalg.qr(rng.standard_normal((n, 40)))
    s = 10.0 * 0.85 ** np.arange(40)
    X = (Z * s) @ W.T + 0.01 * rng.standard_normal((m, n))
    X -= X.mean(axis=0, keepdims=True)
    return np.asfortranarray(X)


# ---- C4: POD snapshot-shaped 50000 x 2000, l = 64 (numerically rank-deficient: the spectrum decays super-geometrically) --
def c4_pod(m: int = 50000, n: int = 2000, seed: int = 4, K: int = 128) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = (np.arange(m) + 0.5) / m
    mu = rng.uniform(0.001, 0.05, n); t = rng.uniform(0.0, 0.05, n)
    k = np.arange(1, K + 1)
    S = np.sin(np.pi * np.outer(x, k))                                    # m x K
    C = (3.0 / k)[:, None] * np.exp(-np.outer(k ** 2, mu * t) * np.pi ** 2)  # K x n
    return np.asfortranarray(S @ C + 1e-10 * rng.standard_normal((m, n)))


def c4_sparse(m: int = 1_000_000, nnz_per_row: int = 10, seed: int = 5):
    """1M x 1M CSR, ~10 uniform random columns per row plus a unit diagonal; int64 rowptr, int32 colidx."""
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, m, size=(m, nnz_per_row), dtype=np.int64)
    cols = np.concatenate([cols, np.arange(m, dtype=np.int64)[:, None]], axis=1)
    vals = np.concatenate([rng.standard_normal((m, nnz_per_row)), np.ones((m, 1))], axis=1)
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1); vals = np.take_along_axis(vals, order, axis=1)
    w = nnz_per_row + 1
    rowptr = np.arange(0, (m + 1) * w, w, dtype=np.int64)
    return rowptr, cols.reshape(-1).astype(np.int32), vals.reshape(-1)


# ---- C5: tall dense m x n FP64, A = X diag(s) Y^T + 1e-6 N, generated shard-by-shard on the device ---------------------
def c5_factors(m: int, n: int, rank: int = 200, seed: int = 6):
    """Small factors that define C5; every rank regenerates them identically."""
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((n, rank)) / np.sqrt(n)
    s = 10.0 ** (-4.0 * np.arange(rank) / rank)
    return Y, s


def c5_shard_torch(m_total: int, n: int, row0: int, rows: int, device, rank: int = 200, seed: int = 6, noise: float = 1e-6):
    """Rows [row0, row0+rows) of C5 as a column-major matrix on `device`: returns a torch tensor t of shape (n, rows)
    (so that t.T is the rows x n matrix and t's storage is its column-major layout).  X rows come from a per-row-block
    device generator, so a shard is the same no matter how the rows are split across ranks (blocks of 1000 rows)."""
    import torch
    Y, s = c5_factors(m_total, n, rank, seed)
    Ys = torch.from_numpy((Y * s).T.copy()).to(device)                    # rank x n
    out = torch.empty((n, rows), dtype=torch.float64, device=device)
    blk = 1000
    b0 = row0 // blk
    r = row0
    while r < row0 + rows:
        b = r // blk
        lo, hi = b * blk, min((b + 1) * blk, m_total)
        g = torch.Generator(device=device); g.manual_seed(seed * 1_000_003 + b)
        Xb = torch.randn((hi - lo, rank), dtype=torch.float64, device=device, generator=g) / (m_total ** 0.5)
        Nb = torch.randn((hi - lo, n), dtype=torch.float64, device=device, generator=g)
        a, e = max(lo, row0), min(hi, row0 + rows)
        blockA = Xb[a - lo:e - lo] @ Ys + noise * Nb[a - lo:e - lo]
        out[:, a - row0:e - row0] = blockA.T
        r = e
    return out


def row_split(m: int, nranks: int, rank: int):
    """rows/P (+1 for the first rows%P ranks): the reference's own split rule (src/rSVD.cpp:20-23)."""
    base, rem = divmod(m, nranks)
    rows = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, rows
