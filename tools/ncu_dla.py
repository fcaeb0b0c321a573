"""one large launch of the DLA loss kernel (for ncu): python tools/ncu_dla.py [B] [L]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ultra_pytorch_b200.engine import RankerEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = RankerEngine(4, [])
s = torch.randn(B, L, device="cuda")
y = (torch.rand(B, L, device="cuda") < 0.2).float()
pw = torch.randn(L, device="cuda") * 0.1
pb = torch.zeros(1, device="cuda")
d = torch.empty(B, L, device="cuda")
dp = torch.zeros(L + 1, device="cuda")
sums = torch.zeros(4, device="cuda")
for i in range(3):
    eng.dla_loss(s, y, pw, pb, d, dp, sums)
torch.cuda.synchronize()
