"""3x3 convs at the 4x4 / 8x8 levels (few pixel tiles) for different forced N tiles (ADM_BN)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit
print("ADM_BN =", os.environ.get("ADM_BN"))
for cin, cout, res in [(384, 384, 4), (768, 384, 4), (384, 384, 8), (768, 384, 8)]:
    x = torch.randn(128, res, res, cin, device="cuda").bfloat16()
    dy = torch.randn(128, res, res, cout, device="cuda").bfloat16()
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 60)
    out = torch.empty(128, res, res, cout, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(128, res, res, cin, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    t = timeit(lambda: ops.conv_fprop(x, w, bias=bias, out=out))
    t2 = timeit(lambda: ops.conv_dgrad(dy, w, out=dx))
    print(f"small [{cin}->{cout} @{res}] fprop {t*1000:6.1f} us  dgrad {t2*1000:6.1f} us")
