"""The pipeline's orthonormalisation step (guarded CholeskyQR2 -> Householder TSQR) on device data:
||Q^T Q - I||, ||Q R - Y|| / ||Y||, which path ran, and the time against the Householder-only policy."""
import json, sys
import torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine

E = Engine(0)
dev = torch.device("cuda:0")
E.set_stream(torch.cuda.current_stream().cuda_stream)
f64 = torch.float64


def make(rows, l, kind):
    g = torch.Generator(device=dev); g.manual_seed(rows * 131 + l)
    Y = torch.randn((l, rows), dtype=f64, device=dev, generator=g)           # column-major rows x l
    if kind == "rank5":
        Y[5:] = torch.randn((l - 5, 5), dtype=f64, device=dev, generator=g) @ Y[:5]
    elif kind.startswith("kappa"):                                            # prescribed condition number 10^e
        e = float(kind[5:])
        Q, _ = torch.linalg.qr(Y.T)                                           # rows x l
        W, _ = torch.linalg.qr(torch.randn((l, l), dtype=f64, device=dev, generator=g))
        s = 10.0 ** (-e * torch.arange(l, dtype=f64, device=dev) / max(l - 1, 1))
        Y = ((Q * s) @ W.T).T.contiguous()
    elif kind == "zero_col":
        Y[l // 2] = 0.0
    return Y


def timed(fn, Y, Y0, reps=5):
    best = 1e30
    for _ in range(reps):
        Y.copy_(Y0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run(rows, l, kind, expect):
    Y0 = make(rows, l, kind); Y = Y0.clone()
    R = torch.zeros((l, l), dtype=f64, device=dev)
    E.set_qr_policy(False)
    path = E.orthonormalize_dev(Y.data_ptr(), rows, l, rows, False, R.data_ptr()); torch.cuda.synchronize()
    Q = Y.T; Rm = R.T
    orth = (Q.T @ Q - torch.eye(l, dtype=f64, device=dev)).norm().item()
    rec = ((Q @ Rm - Y0.T).norm() / Y0.norm()).item()
    tri = torch.tril(Rm, -1).abs().max().item()
    t_auto = timed(lambda: E.orthonormalize_dev(Y.data_ptr(), rows, l, rows, False, None), Y, Y0)
    E.set_qr_policy(True)
    t_hh = timed(lambda: E.orthonormalize_dev(Y.data_ptr(), rows, l, rows, False, None), Y, Y0)
    E.set_qr_policy(False)
    ok = orth < 1e-12 and rec < 1e-13 and tri == 0.0 and (expect is None or path == expect)
    print(json.dumps({"orth": [rows, l, kind], "path": "cholqr2" if path == 0 else "householder", "orth_err": orth, "rec": rec,
                      "ms_auto": round(t_auto, 3), "ms_householder": round(t_hh, 3), "ok": ok}), flush=True)
    return ok


ok = True
for rows, l, kind, expect in [
        (200000, 100, "randn", 0), (25000, 100, "randn", 0), (20000, 100, "randn", 0), (1000000, 64, "randn", 0), (50000, 64, "randn", 0),
        (100000, 20, "randn", 0), (4096, 50, "randn", 0), (256, 100, "randn", 1), (100, 100, "randn", 1), (777, 33, "randn", 0),
        (12345, 37, "kappa3", 0), (20000, 100, "kappa6", 0), (20000, 100, "kappa7", None), (20000, 100, "kappa8", None),
        (20000, 100, "kappa10", 1), (20000, 100, "kappa14", 1), (5000, 50, "rank5", 1), (3000, 64, "zero_col", 1), (2049, 100, "rank5", 1),
        (20001, 101, "randn", 0), (9999, 112, "randn", 0), (30000, 113, "randn", 0), (30000, 128, "randn", 0), (30000, 129, "randn", None), (50000, 200, "randn", None), (8000, 256, "kappa3", None), (6000, 150, "rank5", None), (3000, 96, "randn", 0), (3000, 97, "randn", 0), (3000, 32, "randn", 0), (3000, 65, "randn", 0), (30001, 8, "randn", 1), (5000, 1, "randn", 1), (30001, 16, "randn", 0)]:
    ok &= run(rows, l, kind, expect)
print(json.dumps({"all_ok": bool(ok), "counts": E.qr_path_counts()}))
