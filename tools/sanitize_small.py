"""Smallest run that touches every kernel family (for compute-sanitizer): leaves + cluster nodes + wide panel, Jacobi (both),
power, SpMM, GEMMs with split-K."""
import sys
import numpy as np
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
E = Engine(0)
rng = np.random.default_rng(0)
A = rng.standard_normal((1500, 300))
for l, meth in [(100, SVDMethod.Jacobi), (24, SVDMethod.ParallelJacobi), (8, SVDMethod.Power), (120, SVDMethod.Jacobi)]:
    U, S, V = E.rSVD(A, l, meth, Omega=W.omega(300, l), q=1)
    print(l, int(meth), float(S[0]))
Q, R = E.qr_decomposition_reduced(A[:, :40]); print("qr", float(np.abs(Q.T @ Q - np.eye(40)).max()))
import scipy.sparse as sp
B = sp.random(900, 700, density=0.01, random_state=1, format="csr"); B.sort_indices()
U, S, V = E.rSVD_csr(B.indptr, B.indices, B.data, B.shape, 16, SVDMethod.Jacobi, Omega=W.omega(700, 16), q=1); print("csr", float(S[0]))
E.close(); print("done")
