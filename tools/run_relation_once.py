"""Runs the relation-layer tail kernels (rel_gn forward / backward, bilinear backward) a few times at the DIV2K config's
full-resolution shape over rotating buffers, for `ncu --set full -k regex:"rel_gn|lerp_axis"`.
Usage: python tools/run_relation_once.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hw, c, hq = 128, 128, 16
xs = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(3)]
ys = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(3)]
ds = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(3)]
z = torch.randn(B, hq, hq, c, device="cuda")
gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
for i in range(3):  # per iteration: rel_gn_stats, rel_gn_apply, rel_gn_bwd_sums, rel_gn_bwd_apply, lerp_axis_bwd x 2
    _, stats = ops.rel_gn_fwd(xs[i], ys[i], z, gamma, beta, 8)
    ops.rel_gn_bwd(ds[i], xs[i], ys[i], gamma, stats, 8)
    ops.bilinear_bwd(ds[i], (hq, hq))
torch.cuda.synchronize()
print("done")
