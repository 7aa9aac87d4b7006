"""One orthonormalisation of a rows x l sketch under the default policy (for ncu launch lists)."""
import sys, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
l = int(sys.argv[2]) if len(sys.argv) > 2 else 100
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
Y0 = torch.randn((l, rows), dtype=torch.float64, device=dev); Y = Y0.clone()
for _ in range(3):
    Y.copy_(Y0); E.orthonormalize_dev(Y.data_ptr(), rows, l, rows, False, None)
torch.cuda.synchronize()
print("path counts", E.qr_path_counts())
