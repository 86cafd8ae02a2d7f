import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adcc_b200 import backend as be
n = 780 * 79800
vecs = [torch.randn(n, device="cuda", dtype=torch.float64) for _ in range(9)]
for _ in range(3):
    be.dot_many(vecs[8], vecs[:8])
torch.cuda.synchronize()
print("ok")
