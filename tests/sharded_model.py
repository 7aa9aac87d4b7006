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
