"""One launch of each PCA streaming kernel on a 100000 x 1000 matrix (for `ncu --set full`)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
m, n = 100000, 1000
A = torch.randn((n, m), dtype=torch.float64, device=dev) * 3 + 1.5
mu = torch.empty(n, dtype=torch.float64, device=dev); sd = torch.empty(n, dtype=torch.float64, device=dev)
for _ in range(2):
    E._check(E.lib.rsvdb_column_stats_dev(E.h, A.data_ptr(), m, n, m, mu.data_ptr(), sd.data_ptr()))
    E._check(E.lib.rsvdb_center_columns_dev(E.h, A.data_ptr(), m, n, m, mu.data_ptr(), sd.data_ptr()))
torch.cuda.synchronize()
print("ok", float(mu.abs().max()), float(sd.mean()))
