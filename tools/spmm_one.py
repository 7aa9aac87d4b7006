import sys, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, workloads as W
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
m = 1_000_000; l = 64
rowptr, col, val = W.c4_sparse(m, 10)
rp = torch.from_numpy(rowptr).to(dev); ci = torch.from_numpy(col).to(dev); va = torch.from_numpy(val).to(dev)
X = torch.randn((m, l), dtype=torch.float64, device=dev); Y = torch.empty((m, l), dtype=torch.float64, device=dev)
for _ in range(2):
    assert E.lib.rsvdb_csr_spmm_dev(E.h, m, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), X.data_ptr(), l, Y.data_ptr()) == 0
torch.cuda.synchronize(); print("done")
