"""A few launches of the 3x3 conv with / without the GroupNorm prologue (384 -> 384 @16x16, batch 128) for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops

torch.manual_seed(0)
n, hw, c = 128, 16, 384
xs = [torch.randn(n, hw, hw, c, device="cuda").bfloat16() for _ in range(6)]
wpk = ops.pack_conv_weight(torch.randn(c, c, 3, 3, device="cuda") / 60)
bias = torch.zeros(c, device="cuda")
gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
coef, _ = ops.gn_forward(xs[0], None, gamma, beta, 32, 1e-5, act=True, apply=False)
for i in range(6):
    ops.conv_fprop(xs[i], wpk, bias=bias)
    ops.conv_fprop_gn(xs[i], wpk, coef, bias=bias, act=True, want_act=False)
torch.cuda.synchronize()
print("ok")
