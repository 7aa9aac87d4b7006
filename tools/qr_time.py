import sys, json, torch
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine
E = Engine(0); dev = torch.device("cuda:0"); E.set_stream(torch.cuda.current_stream().cuda_stream)
for rows, l in [(256, 100), (25000, 100), (20000, 100), (200000, 100), (1000000, 64), (50000, 64), (100000, 20), (4096, 50)]:
    Y0 = torch.randn((l, rows), dtype=torch.float64, device=dev); Y = Y0.clone()
    best = 1e30
    for _ in range(5):
        Y.copy_(Y0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); E.qr_dev(Y.data_ptr(), rows, l, rows, False, None); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"rows": rows, "l": l, "ms": round(best, 4)}))
