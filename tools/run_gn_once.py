"""Runs each GroupNorm-family kernel a few times on the dominant CIFAR shape (for `ncu --set full -k regex:gn_`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
dev = "cuda"
n, hw, c = 128, 32, 192
xs = [torch.randn(n, hw, hw, c, device=dev).bfloat16() for _ in range(4)]
dys = [torch.randn(n, hw, hw, c, device=dev).bfloat16() for _ in range(4)]
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
params = torch.randn(n, 2 * c, device=dev) * 0.1
dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
dp = torch.zeros(n, 2 * c, device=dev)
for i in range(3):
    coef = ops.gn_stats(xs[i], None, gamma, beta, 32, 1e-5, params=params)
    y = ops.gn_apply(xs[i], None, coef, act=True, drop_p=0.1, seed=i)
    ops.gn_bwd(dys[i], xs[i], None, coef, gamma, beta, 32, params=params, act=True, drop_p=0.1, seed=i, dgamma=dg,
               dbeta=db, dparams=dp, add=dys[i + 1])
torch.cuda.synchronize()
print("done")
