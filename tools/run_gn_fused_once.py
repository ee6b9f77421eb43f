"""Runs the fused (cluster-per-sample) GroupNorm forward / backward kernels a few times on the dominant CIFAR shape over
rotating buffers, the way UNetEngine.block_fwd / block_bwd call them (`ncu --set full -k regex:gn_.*fused`)."""
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
dg, db, dbias = torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev)
dp = torch.zeros(n, 2 * c, device=dev)
for i in range(3):
    coef, y = ops.gn_forward(xs[i], None, gamma, beta, 32, 1e-5, params=params, act=True, drop_p=0.1, seed=i)
    ops.gn_bwd(dys[i], xs[i], None, coef, gamma, beta, 32, params=params, act=True, drop_p=0.1, seed=i, dgamma=dg,
               dbeta=db, dparams=dp, dbias1=dbias)
torch.cuda.synchronize()
print("done")
