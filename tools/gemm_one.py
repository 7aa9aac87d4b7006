import sys, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
m, n, l = 200000, 20000, 100
A = torch.randn((n, m), dtype=torch.float64, device=dev); X = torch.randn((l, n), dtype=torch.float64, device=dev)
Y = torch.empty((l, m), dtype=torch.float64, device=dev); Q = torch.randn((l, m), dtype=torch.float64, device=dev); Z = torch.empty((l, n), dtype=torch.float64, device=dev)
for _ in range(2):
    E.gemm_an_dev(A.data_ptr(), m, n, m, X.data_ptr(), n, l, Y.data_ptr(), m)
    E.gemm_at_dev(A.data_ptr(), m, n, m, Q.data_ptr(), m, l, Z.data_ptr(), n, False)
torch.cuda.synchronize(); print("done")
