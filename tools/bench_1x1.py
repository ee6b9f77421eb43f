"""Epilogue-bound shapes: 1x1 convs of the CIFAR net (fprop) in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit
for cin, cout, res in [(384, 1152, 16), (384, 1152, 8), (384, 384, 16), (768, 384, 16), (384, 192, 32)]:
    x = torch.randn(128, res, res, cin, device="cuda").bfloat16()
    w = ops.pack_conv_weight(torch.randn(cout, cin, 1, 1, device="cuda") / 20)
    out = torch.empty(128, res, res, cout, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    t = timeit(lambda: ops.conv_fprop(x, w, bias=bias, out=out))
    fl = 2.0 * 128 * res * res * cin * cout
    mb = 128 * res * res * (cin + cout) * 2 / 1e6
    print(f"1x1 [{cin}->{cout} @{res}] {t*1000:7.1f} us  {fl/t/1e9:7.1f} TF/s  {mb/t/1e3:7.1f} GB/s (in+out)")
