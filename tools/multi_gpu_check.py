"""Multi-GPU parity: run under torchrun with N ranks; every rank holds a row shard; compare with the single-process oracle."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
from oracle import rsvd_oracle as O

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
E = Engine(local)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(E.comm_unique_id()), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
E.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
E.set_stream(torch.cuda.current_stream().cuda_stream)
ok_all = True
for (m, n, l, q, gen) in [(3001, 400, 32, 2, "pod"), (20000, 1500, 100, 2, "gauss"), (1000, 300, 16, 1, "rank2"), (50000, 2000, 64, 2, "pod"),
                          (3000, 500, 120, 1, "gauss"), (2001, 640, 101, 2, "gauss"), (12000, 700, 160, 1, "gauss")]:     # wide / odd sketches: block Gram-Schmidt TSQR over shards
    rng = np.random.default_rng(42)
    if gen == "pod": A = W.c4_pod(m, n)
    elif gen == "rank2": A = np.asfortranarray(np.outer(rng.standard_normal(m), rng.standard_normal(n)) + np.outer(rng.standard_normal(m), rng.standard_normal(n)))
    else: A = np.asfortranarray(rng.standard_normal((m, 60)) @ np.diag(0.9 ** np.arange(60)) @ rng.standard_normal((60, n)) + 1e-3 * rng.standard_normal((m, n)))
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    # host API on the local shard (every rank calls collectively)
    U_p, S, V = E.rSVD(A[off:off + rows], l, SVDMethod.Jacobi, Omega=Om, q=q)
    Ut = torch.from_numpy(np.ascontiguousarray(U_p)).to(dev)
    parts = [torch.empty((W.row_split(m, world, r)[1], U_p.shape[1]), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(parts, Ut)
    U = torch.cat(parts, 0).cpu().numpy()
    Sall = [torch.empty(len(S), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(Sall, torch.from_numpy(S).to(dev))
    same_S = all(torch.equal(Sall[0], s) for s in Sall)
    # the guard of the CholeskyQR2 fast path must send every rank down the same branch (its collectives differ from the TSQR's)
    cnt = torch.tensor(E.qr_path_counts(), dtype=torch.int64, device=dev)
    call = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(call, cnt)
    same_S = same_S and all(torch.equal(call[0], x) for x in call)
    if rank == 0:
        Uo, So, Vo = O.rsvd(A, Om, l, q, O.JACOBI)
        okS, relS = O.sigma_close(S, So)
        eg, eo = O.reconstruction_error(A, U, S, V), O.reconstruction_error(A, Uo, So, Vo)
        orthU = np.linalg.norm(U.T @ U - np.eye(U.shape[1]))
        ok = okS and eg <= eo + 1e-8 * np.linalg.norm(A) and orthU < 1e-10 and same_S
        ok_all &= ok
        print(json.dumps({"multi_gpu": [m, n, l, q, gen], "world": world, "ok": bool(ok), "relS": relS, "err_gpu": eg, "err_oracle": eo, "orthU": orthU, "S_and_qr_paths_identical_on_all_ranks": same_S,
                          "qr_paths_cholqr2_householder_so_far": [int(x) for x in cnt.tolist()]}), flush=True)
# the ADVICE round-1 edge: shard heights straddle the wide-panel threshold 4*l (shards of 480,...,479 rows at P = 4, l = 120): every rank
# must take the same QR branch or the collectives mismatch (pipeline.cu qr_inplace agrees on the shortest shard first)
for l in (120,):
    m, n, q = 4 * world * l - 1, 300, 1
    rng = np.random.default_rng(77)
    A = np.asfortranarray(rng.standard_normal((m, 40)) @ rng.standard_normal((40, n)) + 1e-3 * rng.standard_normal((m, n)))
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    U_p, S, V = E.rSVD(A[off:off + rows], l, SVDMethod.Jacobi, Omega=Om, q=q)
    if rank == 0:
        Uo, So, Vo = O.rsvd(A, Om, l, q, O.JACOBI)
        okS, relS = O.sigma_close(S, So)
        ok_all &= okS
        print(json.dumps({"multi_gpu_wide_threshold": [m, n, l, q], "shard_rows": [W.row_split(m, world, r)[1] for r in range(world)], "world": world, "ok": bool(okS), "relS": relS}), flush=True)
# row-sharded CSR (north_star: sparse .mtx inputs): every rank holds a row block of the CSR matrix, A^T Q partial sums are all-reduced
import scipy.sparse as sp
for (m, n, l, q, dens) in [(6001, 900, 24, 2, 0.01), (20000, 20000, 32, 1, 0.0008)]:
    M = sp.random(m, n, density=dens, format="csr", random_state=np.random.default_rng(m), data_rvs=np.random.default_rng(m + 1).standard_normal) + sp.eye(m, n, format="csr")
    M = M.tocsr(); M.sort_indices()
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    Mp = M[off:off + rows]
    U_p, S, V = E.rSVD_csr(Mp.indptr.astype(np.int64), Mp.indices.astype(np.int32), Mp.data.astype(np.float64), (rows, n), l, SVDMethod.Jacobi, Omega=Om, q=q)
    Ut = torch.from_numpy(np.ascontiguousarray(U_p)).to(dev)
    parts = [torch.empty((W.row_split(m, world, r)[1], U_p.shape[1]), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(parts, Ut)
    U = torch.cat(parts, 0).cpu().numpy()
    if rank == 0:
        A = np.asfortranarray(M.toarray())
        Uo, So, Vo = O.rsvd(A, Om, l, q, O.JACOBI)
        okS, relS = O.sigma_close(S, So)
        eg, eo = O.reconstruction_error(A, U, S, V), O.reconstruction_error(A, Uo, So, Vo)
        orthU = np.linalg.norm(U.T @ U - np.eye(U.shape[1]))
        ok = okS and eg <= eo + 1e-8 * np.linalg.norm(A) and orthU < 1e-10
        ok_all &= ok
        print(json.dumps({"multi_gpu_csr": [m, n, l, q, int(M.nnz)], "world": world, "ok": bool(ok), "relS": relS, "err_gpu": eg, "err_oracle": eo, "orthU": orthU}), flush=True)
# randomized PCA on row shards: column statistics are all-reduced, the centring corrections are applied per shard
for (m, n, l, normalize) in [(20001, 300, 20, True), (6000, 500, 32, False)]:
    rng = np.random.default_rng(5)
    A = W.c3_pca(m, n, seed=9) * (1.0 + (np.arange(n) % 7)) + 4.0 * rng.standard_normal(n)
    Om = W.omega(n, l)
    off, rows = W.row_split(m, world, rank)
    mean, sd, U_p, S, V = E.rpca(A[off:off + rows], l, normalize, SVDMethod.Jacobi, Om, 2)
    Ut = torch.from_numpy(np.ascontiguousarray(U_p)).to(dev)
    parts = [torch.empty((W.row_split(m, world, r)[1], U_p.shape[1]), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(parts, Ut)
    U = torch.cat(parts, 0).cpu().numpy()
    if rank == 0:
        mu = A.sum(axis=0) / m; C = A - mu
        ok = bool(np.allclose(mean, mu, rtol=1e-12, atol=1e-12))
        if normalize:
            s = np.sqrt((C * C).sum(axis=0) / (m - 1)); C = C / s
            ok &= bool(np.allclose(sd, s, rtol=1e-12))
        Uo, So, Vo = O.rsvd(np.asfortranarray(C), Om, l, 2, O.JACOBI)
        okS, relS = O.sigma_close(S, So)
        eg, eo = O.reconstruction_error(C, U, S, V), O.reconstruction_error(C, Uo, So, Vo)
        ok &= okS and eg <= eo + 1e-8 * np.linalg.norm(C)
        ok_all &= ok
        print(json.dumps({"multi_gpu_rpca": [m, n, l, normalize], "world": world, "ok": bool(ok), "relS": relS, "err_gpu": eg, "err_oracle": eo}), flush=True)
if rank == 0:
    print(json.dumps({"all_ok": bool(ok_all)}), flush=True)
E.close()
dist.destroy_process_group()
