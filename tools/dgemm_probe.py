"""cuBLAS FP64 GEMM probe (library denominator for the FP64 roofline).  Prints JSON lines."""
import json, torch
def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
dev = torch.device("cuda:0")
for (m, n, k) in [(8192, 8192, 8192), (65536, 104, 8192), (104, 8192, 65536), (131072, 128, 16384), (65536, 16, 8192), (65536, 64, 8192)]:
    a = torch.randn(m, k, device=dev, dtype=torch.float64)
    b = torch.randn(k, n, device=dev, dtype=torch.float64)
    ms = t(lambda: torch.matmul(a, b))
    print(json.dumps({"probe": "cublas_dgemm", "m": m, "n": n, "k": k, "ms": round(ms, 4), "tflops": round(2.0 * m * n * k / ms * 1e-9, 3)}), flush=True)
    del a, b
# sustained 3 s
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
import time
torch.cuda.synchronize(); t0 = time.time(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while n < 100:
    torch.matmul(a, b); n += 1
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"probe": "cublas_dgemm_sustained", "seconds": round(ms * 1e-3, 2), "tflops": round(n * 2.0 * 8192**3 / ms * 1e-9, 3)}), flush=True)
