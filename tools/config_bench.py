"""Device-resident timing of BASELINE.json configs 1-4 (config 5 is bench.py): ms per rSVD, algorithmic GFLOP/s and HBM GB/s,
phase breakdown, and the CPU oracle on the same input."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
from oracle import rsvd_oracle as O
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
cases = [("c1_ramp_100x100_l16", W.c1_ramp(100), 16), ("c1_identity_160_l16", W.c1_identity(160), 16), ("c2_image_4096x4096_l50", W.c2_image(), 50),
         ("c3_pca_100000x1000_l20", W.c3_pca(), 20), ("c4_pod_50000x2000_l64", W.c4_pod(), 64)]
for name, A, l in cases:
    m, n = A.shape; q = 2
    Om = W.omega(n, l)
    Ad = torch.from_numpy(np.ascontiguousarray(A.T)).to(dev); Od = torch.from_numpy(np.ascontiguousarray(Om.T)).to(dev)
    U = torch.empty((l, m), dtype=torch.float64, device=dev); V = torch.empty((l, n), dtype=torch.float64, device=dev); S = torch.empty(l, dtype=torch.float64, device=dev)
    run = lambda: E.rsvd_dev(Ad.data_ptr(), m, n, m, Od.data_ptr(), n, l, q, SVDMethod.Jacobi, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n)
    for _ in range(3): run()
    torch.cuda.synchronize(); E.set_profiling(True); E.phase_ms()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps; ph = {k: round(v / reps, 4) for k, v in E.phase_ms().items()}; E.set_profiling(False)
    t0 = time.perf_counter(); Uo, So, Vo = O.rsvd(A, Om, l, q, O.JACOBI); cpu_ms = (time.perf_counter() - t0) * 1e3
    ok, rel = O.sigma_close(S.cpu().numpy(), So)
    F = 12.0 * m * n * l; B = 6 * 8.0 * m * n
    print(json.dumps({"config": name, "ms": round(ms, 4), "gflops": round(F / ms * 1e-6, 1), "hbm_GBps_A_stream": round(B / ms * 1e-6, 1), "phases_ms": ph,
                      "cpu_oracle_ms": round(cpu_ms, 1), "speedup_vs_cpu_oracle": round(cpu_ms / ms, 1), "sigma_parity": ok, "max_rel_sigma_err": rel}), flush=True)
