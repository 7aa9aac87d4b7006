import sys, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
rows, l = int(sys.argv[1]) if len(sys.argv) > 1 else 37888, int(sys.argv[2]) if len(sys.argv) > 2 else 100
Y = torch.randn((l, rows), dtype=torch.float64, device=dev)
for _ in range(2):
    E.qr_dev(Y.data_ptr(), rows, l, rows, False, None)
torch.cuda.synchronize()
print("done")
