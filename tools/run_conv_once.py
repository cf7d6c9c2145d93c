"""A few launches of the res5 3x3 convolution (R = 1000 RoIs, 512 -> 512 channels, 7x7) under one TF32 pass and under 3xTF32
(for ncu captures of conv_tc2d_kernel / conv_tc2_kernel in 3x3 mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
x = torch.randn(1000, 7, 7, 512, generator=g).to(dev).permute(0, 3, 1, 2)
taps = ops.conv_taps((torch.randn(512, 512, 3, 3, generator=g) / 68.0).to(dev))
split = ops.conv_split_weights(taps)
b = torch.randn(512, generator=g).to(dev)
for _ in range(3):
    ops.conv3x3(x, taps, b, relu=True, precision="tf32")
    ops.conv3x3(x, taps, b, relu=True, precision="fp32", w_split=split)
torch.cuda.synchronize()
print("ok")
