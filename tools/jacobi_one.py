import sys, numpy as np
sys.path.insert(0, ".")
from rsvd_kamaneh_raganato_terrana_b200 import Engine, SVDMethod
E = Engine(0)
B = np.asfortranarray(np.random.default_rng(1).standard_normal((100, 100)))
for _ in range(2):
    E.svd(B, SVDMethod.Jacobi)
print("done", E.last_svd_info())
