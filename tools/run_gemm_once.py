"""A few launches of the relation contraction at the cfg3 shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
a = torch.randn(49000, 256, generator=g).to(dev)
w = (torch.randn(256, 512, generator=g) / 16).to(dev)
for _ in range(4):
    ops.gemm_nt(a, w[:, :256], None, "fp32")
torch.cuda.synchronize()
print("ok")
