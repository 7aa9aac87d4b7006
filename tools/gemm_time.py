import sys, json, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for (m, n, l) in [(200000, 20000, 100), (25000, 20000, 100), (50000, 2000, 64)]:
    A = torch.randn((n, m), dtype=torch.float64, device=dev); X = torch.randn((l, n), dtype=torch.float64, device=dev)
    Y = torch.empty((l, m), dtype=torch.float64, device=dev); Q = torch.randn((l, m), dtype=torch.float64, device=dev); Z = torch.empty((l, n), dtype=torch.float64, device=dev)
    an = t(lambda: E.gemm_an_dev(A.data_ptr(), m, n, m, X.data_ptr(), n, l, Y.data_ptr(), m))
    at = t(lambda: E.gemm_at_dev(A.data_ptr(), m, n, m, Q.data_ptr(), m, l, Z.data_ptr(), n, False))
    fl = 2.0 * m * n * l
    print(json.dumps({"shape": [m, n, l], "an_ms": round(an, 3), "an_tf": round(fl / an * 1e-9, 2), "at_ms": round(at, 3), "at_tf": round(fl / at * 1e-9, 2)}), flush=True)
    del A, X, Y, Q, Z
