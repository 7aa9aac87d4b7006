"""K7 timing: CSR SpMM (1M x 1M, ~11 nnz/row, l = 64) against the HBM roofline; sparse rSVD phases."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
m = 1_000_000; l = 64
rowptr, col, val = W.c4_sparse(m, 10); nnz = int(rowptr[-1])
rp = torch.from_numpy(rowptr).to(dev); ci = torch.from_numpy(col).to(dev); va = torch.from_numpy(val).to(dev)
X = torch.randn((m, l), dtype=torch.float64, device=dev); Y = torch.empty((m, l), dtype=torch.float64, device=dev)
def run(): 
    rc = E.lib.rsvdb_csr_spmm_dev(E.h, m, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), X.data_ptr(), l, Y.data_ptr()); assert rc == 0
run(); run(); torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
comp = 12.0 * nnz + 8.0 * (m + 1) + 8.0 * l * (m + m); gath = 12.0 * nnz + 8.0 * (m + 1) + 8.0 * l * nnz + 8.0 * l * m
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6549.1}
print(json.dumps({"spmm": [m, m, nnz, l], "ms": round(best, 4), "compulsory_GBps": round(comp / best * 1e-6, 1), "gather_inclusive_GBps": round(gath / best * 1e-6, 1),
                  "gflops": round(2.0 * nnz * l / best * 1e-6, 1), "hbm_peak_GBps": peaks["hbm_gbs"], "frac_compulsory": round(comp / best * 1e-6 / peaks["hbm_gbs"], 3),
                  "frac_gather_inclusive": round(gath / best * 1e-6 / peaks["hbm_gbs"], 3)}), flush=True)
del X, Y
E.lib.rsvdb_use_own_stream(E.h)
E.set_profiling(True)
E.rSVD_csr(rowptr, col, val, (m, m), l, SVDMethod.Jacobi, Omega=None, q=2, seed=1); E.phase_ms()
t0 = time.time(); U, S, V = E.rSVD_csr(rowptr, col, val, (m, m), l, SVDMethod.Jacobi, Omega=None, q=2, seed=1); dt = time.time() - t0
print(json.dumps({"rsvd_csr_1Mx1M_l64_q2": {"e2e_ms": round(dt * 1e3, 1), "phases_ms": {k: round(v, 3) for k, v in E.phase_ms().items()}, "sigma_head": [float(x) for x in S[:3]]}}), flush=True)
