"""Fused GroupNorm kernels on the step's shapes (batch 128), isolated: us per launch and algorithmic TB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit

torch.manual_seed(0)
for (hw, c1, c2) in [(32, 192, 0), (32, 384, 192), (32, 192, 192), (16, 384, 0), (16, 384, 384), (16, 384, 192), (8, 384, 0), (8, 384, 384), (4, 384, 384)]:
    n, C = 128, c1 + c2
    xs = [(torch.randn(n, hw, hw, c1, device="cuda").bfloat16(), torch.randn(n, hw, hw, c2, device="cuda").bfloat16() if c2 else None)
          for _ in range(3)]
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
    dys = [torch.randn(n, hw, hw, C, device="cuda").bfloat16() for _ in range(3)]
    i = [0]

    def fwd():
        x1, x2 = xs[i[0] % 3]; i[0] += 1
        return ops.gn_forward(x1, x2, gamma, beta, min(32, C // 4), 1e-5, act=True, drop_p=0.1, seed=5)
    coef, _ = fwd()
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")

    def bwd():
        x1, x2 = xs[i[0] % 3]; dy = dys[i[0] % 3]; i[0] += 1
        return ops.gn_bwd(dy, x1, x2, coef, gamma, beta, min(32, C // 4), act=True, drop_p=0.1, seed=5, dgamma=dg, dbeta=db,
                          dy_scratch=True)
    tf, tb = timeit(fwd), timeit(bwd)
    el = n * hw * hw * C
    print(f"GN [128,{hw},{hw},{c1}+{c2}]: fwd {tf*1000:7.1f} us ({6 * el / tf / 1e9:5.2f} TB/s of 6 B/elem)  bwd {tb*1000:7.1f} us "
          f"({10 * el / tb / 1e9:5.2f} TB/s of 10 B/elem)", flush=True)
