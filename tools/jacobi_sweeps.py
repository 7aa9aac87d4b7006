"""Sweeps / rotations of the small Jacobi SVD inside the rSVD for the BASELINE configs (scaled where needed)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod, workloads as W
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
def c5_small(m, n):
    g = torch.Generator(device=dev); g.manual_seed(6)
    X = torch.randn((m, 200), dtype=torch.float64, device=dev, generator=g) / np.sqrt(m); Y = torch.randn((n, 200), dtype=torch.float64, device=dev, generator=g) / np.sqrt(n)
    s = 10.0 ** (-4.0 * torch.arange(200, dtype=torch.float64, device=dev) / 200)
    return ((X * s) @ Y.T + 1e-6 * torch.randn((m, n), dtype=torch.float64, device=dev, generator=g)).T.contiguous()   # stored n x m = column-major m x n
cases = [("c5_like_40000x20000_l100", None, 40000, 20000, 100), ("c2_image_l50", W.c2_image(), 4096, 4096, 50), ("c4_pod_l64", W.c4_pod(), 50000, 2000, 64),
         ("c3_pca_l20", W.c3_pca(), 100000, 1000, 20)]
for name, A, m, n, l in cases:
    Ad = c5_small(m, n) if A is None else torch.from_numpy(np.ascontiguousarray(A.T)).to(dev)
    Od = torch.from_numpy(np.ascontiguousarray(W.omega(n, l).T)).to(dev)
    U = torch.empty((l, m), dtype=torch.float64, device=dev); V = torch.empty((l, n), dtype=torch.float64, device=dev); S = torch.empty(l, dtype=torch.float64, device=dev)
    E.rsvd_dev(Ad.data_ptr(), m, n, m, Od.data_ptr(), n, l, 2, SVDMethod.Jacobi, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n)
    torch.cuda.synchronize()
    sw, rot = E.last_svd_info()
    E.set_profiling(True); E.phase_ms()
    E.rsvd_dev(Ad.data_ptr(), m, n, m, Od.data_ptr(), n, l, 2, SVDMethod.Jacobi, U.data_ptr(), m, S.data_ptr(), V.data_ptr(), n)
    torch.cuda.synchronize(); ph = E.phase_ms(); E.set_profiling(False)
    print(json.dumps({"case": name, "sweeps": sw, "rotations": rot, "pairs_per_sweep": l * (l - 1) // 2, "small_svd_ms": round(ph["small_svd"], 3),
                      "sigma_ratio_last_first": float(S[-1] / S[0])}), flush=True)
