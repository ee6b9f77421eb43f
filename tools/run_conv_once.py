"""Runs conv fprop / dgrad / wgrad of the dominant CIFAR shapes a few times over rotating (cold) buffers, for
`ncu --set full -k regex:tc_gemm`.  Usage: python tools/run_conv_once.py [cin cout res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
cin, cout, res = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (384, 384, 16)
n = 128
nbuf = 6
xs = [torch.randn(n, res, res, cin, device="cuda").bfloat16() for _ in range(nbuf)]
dys = [torch.randn(n, res, res, cout, device="cuda").bfloat16() for _ in range(nbuf)]
w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 60)
bias = torch.zeros(cout, device="cuda")
dw = torch.zeros(cout, 9, cin, device="cuda")
for i in range(3):
    ops.conv_fprop(xs[i], w, bias=bias)
    ops.conv_dgrad(dys[i], w)
    ops.conv_wgrad(dys[i], xs[i], ntaps=9, out=dw)
torch.cuda.synchronize()
print("done")
