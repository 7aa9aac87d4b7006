"""TEST INFRASTRUCTURE: numpy + torch.distributed model of the row-sharded rSVD, rank for rank the communication pattern
of rsvd_kamaneh_raganato_terrana_b200/csrc/pipeline.cu (qr_inplace's sharded branch and gemm_at_phase's all-reduce)."""
import numpy as np


def _allreduce(x, dist, torch):
    t = torch.from_numpy(np.ascontiguousarray(x)); dist.all_reduce(t); return t.numpy()


def _allgather(x, dist, torch):
    t = torch.from_numpy(np.ascontiguousarray(x)); out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t); return [o.numpy() for o in out]


def tsqr_sharded(Y_p, dist, torch, O):
    """local Householder QR -> all-gather R_p -> QR of the stacked R's on every rank -> Q_p = Q_local * Q_stack[p block]."""
    l = Y_p.shape[1]
    Q1, R1 = O.householder_qr(Y_p)
    stack = np.vstack(_allgather(R1, dist, torch))
    Q2, R = O.householder_qr(stack)
    p = dist.get_rank()
    return Q1 @ Q2[p * l:(p + 1) * l, :], R


CHOL_DEV_TOL = 0.05


def _chol_inv(G):
    """cholqr.cu k_chol_inv: (breakdown, ||G - I||_F^2, X = R^-1, R) of G = R^T R; breakdown = first non-positive pivot."""
    l = G.shape[0]
    dev2 = float(np.sum((np.tril(G) - np.eye(l)) ** 2) + np.sum(np.tril(G, -1) ** 2))
    Lw = np.tril(G).copy(); Wi = np.eye(l)
    for j in range(l):
        d = Lw[j, j]
        if not (d > 0.0) or not np.isfinite(d):
            return j + 1, dev2, None, None
        s = 1.0 / np.sqrt(d)
        Lw[j:, j] *= s; Wi[j, :j + 1] *= s
        for i in range(j + 1, l):
            Lw[i, j + 1:i + 1] -= Lw[i, j] * Lw[j + 1:i + 1, j]
            Wi[i, :j + 1] -= Lw[i, j] * Wi[j, :j + 1]
    return 0, dev2, Wi.T, Lw.T


def orthonormalize_sharded(Y_p, dist, torch, O, state):
    """cholqr.cu orthonormalize on a row-sharded sketch: Gram matrices are all-reduced (every rank factors the same bits and takes
    the same decision), the guard reads the SECOND Gram matrix before Y is overwritten, the TSQR is the fallback."""
    if not state.get("failed", False):
        bad1, _, X1, R1 = _chol_inv(_allreduce(Y_p.T @ Y_p, dist, torch))
        if not bad1:
            T_p = Y_p @ X1
            bad2, dev2, X2, R2 = _chol_inv(_allreduce(T_p.T @ T_p, dist, torch))
            if not bad2 and dev2 <= CHOL_DEV_TOL ** 2:
                state["fast"] = state.get("fast", 0) + 1
                return T_p @ X2, R2 @ R1
        state["failed"] = True
    state["householder"] = state.get("householder", 0) + 1
    return tsqr_sharded(Y_p, dist, torch, O)


def orthonormalize_replicated(Z, O, state):
    if not state.get("failed", False):
        bad1, _, X1, R1 = _chol_inv(Z.T @ Z)
        if not bad1:
            T = Z @ X1
            bad2, dev2, X2, R2 = _chol_inv(T.T @ T)
            if not bad2 and dev2 <= CHOL_DEV_TOL ** 2:
                state["fast"] = state.get("fast", 0) + 1
                return T @ X2, R2 @ R1
        state["failed"] = True
    state["householder"] = state.get("householder", 0) + 1
    return O.householder_qr(Z)


def rsvd_sharded_guarded(A_p, Omega, l, q, dist, torch, O, state):
    """rsvd_sharded with the orthonormalisations the product runs by default (guarded CholeskyQR2, Householder fallback)."""
    state.clear()
    Q_p, _ = orthonormalize_sharded(A_p @ Omega, dist, torch, O, state)
    for _ in range(q):
        Qz, _ = orthonormalize_replicated(_allreduce(A_p.T @ Q_p, dist, torch), O, state)
        Q_p, _ = orthonormalize_sharded(A_p @ Qz, dist, torch, O, state)
    Bt = _allreduce(A_p.T @ Q_p, dist, torch)
    Qb, Rb = orthonormalize_replicated(Bt, O, state)                    # include/SVD_class.hpp:116-123: QR of B^T, Jacobi on R^T
    Ur, S, Vr, _ = O.svd_jacobi(Rb.T)
    return Q_p @ Ur, S, Qb @ Vr


def rsvd_sharded(A_p, Omega, l, q, dist, torch, O):
    Q_p, _ = tsqr_sharded(A_p @ Omega, dist, torch, O)                 # shard-local product, distributed QR
    for _ in range(q):
        Z = _allreduce(A_p.T @ Q_p, dist, torch)                        # n x l partial sums
        Qz, _ = O.householder_qr(Z)                                     # replicated
        Q_p, _ = tsqr_sharded(A_p @ Qz, dist, torch, O)
    Bt = _allreduce(A_p.T @ Q_p, dist, torch)                           # B^T = A^T Q
    Ut, S, V, _ = O.svd_jacobi(Bt.T)                                    # replicated small SVD
    return Q_p @ Ut, S, V


def column_stats_sharded(A_p, normalize, dist, torch):
    """pca.cu column_stats: all-reduce [column sums | row count], mean; then all-reduce the centred sums of squares."""
    n = A_p.shape[1]
    tot = _allreduce(np.r_[A_p.sum(axis=0), float(A_p.shape[0])], dist, torch)
    mean = tot[:n] / tot[n]
    inv_sd = None
    if normalize:
        css = _allreduce(((A_p - mean) ** 2).sum(axis=0), dist, torch)
        inv_sd = 1.0 / np.sqrt(css / (tot[n] - 1.0))
    return mean, inv_sd


def rpca_sharded(A_p, Omega, l, q, normalize, dist, torch, O):
    """pipeline.cu with a Centering: the products stream the UNCENTRED shard; (A - 1 mu^T) D X = A (D X) - 1 (mu^T D X) needs no
    exchange, D (A^T - mu 1^T) Q = D (A^T Q - mu (1^T Q)) is corrected per shard BEFORE the all-reduce (it is linear)."""
    mean, inv_sd = column_stats_sharded(A_p, normalize, dist, torch)
    D = np.ones(A_p.shape[1]) if inv_sd is None else inv_sd

    def an(X):
        Xs = X * D[:, None]
        return A_p @ Xs - (mean @ Xs)[None, :]

    def at(Q_p):
        return _allreduce((A_p.T @ Q_p - np.outer(mean, Q_p.sum(axis=0))) * D[:, None], dist, torch)

    Q_p, _ = tsqr_sharded(an(Omega), dist, torch, O)
    for _ in range(q):
        Qz, _ = O.householder_qr(at(Q_p))
        Q_p, _ = tsqr_sharded(an(Qz), dist, torch, O)
    Ut, S, V, _ = O.svd_jacobi(at(Q_p).T)
    return mean, inv_sd, Q_p @ Ut, S, V
