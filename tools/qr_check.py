"""TSQR correctness (||Q^T Q - I||, ||Q R - Y|| / ||Y||) and timing on device data."""
import json, sys
import torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine

E = Engine(0)
dev = torch.device("cuda:0")
E.set_stream(torch.cuda.current_stream().cuda_stream)

def run(rows, l, kind="randn", reps=3):
    Y = torch.randn((l, rows), dtype=torch.float64, device=dev)     # column-major rows x l
    if kind == "rank5":
        Y[5:] = (torch.randn((l - 5, 5), dtype=torch.float64, device=dev) @ Y[:5])
    if kind == "graded":
        Y *= (10.0 ** (-12.0 * torch.arange(l, dtype=torch.float64, device=dev) / l))[:, None]
    Y0 = Y.clone(); R = torch.zeros((l, l), dtype=torch.float64, device=dev)
    E.qr_dev(Y.data_ptr(), rows, l, rows, False, R.data_ptr()); torch.cuda.synchronize()
    Q = Y.T; Rm = R.T                                               # R stored column-major l x l -> R.T is the matrix
    orth = (Q.T @ Q - torch.eye(l, dtype=torch.float64, device=dev)).norm().item()
    rec = ((Q @ Rm - Y0.T).norm() / Y0.norm()).item()
    tri = torch.tril(Rm, -1).abs().max().item()
    best = 1e30
    for _ in range(reps):
        Y.copy_(Y0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); E.qr_dev(Y.data_ptr(), rows, l, rows, False, None); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ok = orth < 1e-12 and rec < 1e-13 and tri == 0.0
    print(json.dumps({"qr": [rows, l, kind], "orth": orth, "rec": rec, "tri": tri, "ms": round(best, 3), "ok": ok}), flush=True)
    return ok

ok = True
for rows, l, kind in [(256, 100, "randn"), (200, 100, "randn"), (1000, 100, "randn"), (777, 33, "randn"), (100, 16, "randn"), (5000, 50, "rank5"), (3000, 64, "graded"),
                      (20000, 100, "randn"), (25000, 100, "randn"), (200000, 100, "randn"), (50000, 64, "randn"), (100000, 20, "randn"), (4096, 50, "randn"),
                      (1025, 100, "randn"), (2048, 64, "randn"), (2049, 100, "rank5"), (32768, 100, "randn"), (32769, 100, "randn"), (12345, 37, "graded"), (30001, 8, "randn"), (20000, 103, "randn"), (20000, 104, "randn"), (20000, 128, "randn"), (50000, 200, "randn"), (6000, 150, "rank5"), (8000, 256, "graded"), (5000, 7, "randn"), (64, 64, "randn"), (120, 100, "randn")]:
    ok &= run(rows, l, kind)
print(json.dumps({"all_ok": bool(ok)}))
