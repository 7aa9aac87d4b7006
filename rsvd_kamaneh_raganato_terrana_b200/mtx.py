"""MatrixMarket IO for the reference's inputs and outputs (`Eigen::loadMarket` / `saveMarket`, reference
tests/rSVD_test.cpp:54-57,113-115): coordinate real general, 1-based, densified on load exactly like the reference
(`MatrixXd(sparseMatrix)`; duplicate entries add up), plus a CSR loader for the sparse path."""
from __future__ import annotations

import numpy as np


def _header(f):
    line = f.readline()
    if not line.startswith("%%MatrixMarket"):
        raise ValueError("not a MatrixMarket file")
    parts = line.split()
    fmt, field, sym = parts[2].lower(), parts[3].lower(), parts[4].lower()
    line = f.readline()
    while line.startswith("%") or not line.strip():
        line = f.readline()
    return fmt, field, sym, [int(x) for x in line.split()]


def load_dense(path) -> np.ndarray:
    with open(path) as f:
        fmt, field, sym, dims = _header(f)
        if fmt == "array":
            m, n = dims[:2]
            vals = np.array(f.read().split(), dtype=np.float64)
            return np.asfortranarray(vals.reshape((m, n), order="F"))
        m, n, nnz = dims
        A = np.zeros((m, n), order="F")
        data = np.array(f.read().split(), dtype=np.float64).reshape(nnz, -1)
        i = data[:, 0].astype(np.int64) - 1; j = data[:, 1].astype(np.int64) - 1
        v = data[:, 2] if data.shape[1] > 2 else np.ones(nnz)
        np.add.at(A, (i, j), v)
        if sym == "symmetric":
            off = i != j
            np.add.at(A, (j[off], i[off]), v[off])
        return A


def load_csr(path):
    """Returns (m, n, rowptr int64, colidx int32, values float64), rows sorted by column."""
    with open(path) as f:
        fmt, field, sym, dims = _header(f)
        if fmt != "coordinate":
            raise ValueError("CSR loading needs a coordinate file")
        m, n, nnz = dims
        data = np.array(f.read().split(), dtype=np.float64).reshape(nnz, -1)
    i = data[:, 0].astype(np.int64) - 1; j = data[:, 1].astype(np.int64) - 1
    v = data[:, 2] if data.shape[1] > 2 else np.ones(nnz)
    order = np.lexsort((j, i))
    i, j, v = i[order], j[order], v[order]
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.add.at(rowptr, i + 1, 1)
    return m, n, np.cumsum(rowptr), j.astype(np.int32), v


def save_coordinate(path, A, tol: float = 0.0):
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    i, j = np.nonzero(np.abs(A) > tol) if tol > 0 else np.nonzero(np.ones_like(A, dtype=bool))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{A.shape[0]} {A.shape[1]} {len(i)}\n")
        for a, b in zip(i, j):
            f.write(f"{a + 1} {b + 1} {A[a, b]:.18e}\n")


def save_array(path, A):
    A = np.asarray(A, dtype=np.float64)
    if A.ndim == 1:
        A = A[:, None]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write(f"{A.shape[0]} {A.shape[1]}\n")
        for v in A.ravel(order="F"):
            f.write(f"{v:.18e}\n")
