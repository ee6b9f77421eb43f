"""3x3 wgrad timing per shape; run with ADM_WGRAD_ROWS=0 / 1 (read at first use) for the per-tap vs tap-row kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit

N = 128
for cin, cout, res in [(192, 192, 32), (384, 192, 32), (384, 384, 16), (768, 384, 16), (384, 384, 8), (576, 192, 32)]:
    flops = 2.0 * N * res * res * cin * cout * 9
    x = torch.randn(N, res, res, cin, device="cuda").bfloat16()
    dy = torch.randn(N, res, res, cout, device="cuda").bfloat16()
    kpad = (cin + 63) // 64 * 64
    dw = torch.zeros(cout, 9 * kpad, device="cuda")
    t = timeit(lambda: ops.conv_wgrad(dy, x, ntaps=9, out=dw))
    print(f"rows={os.environ.get('ADM_WGRAD_ROWS', '1')} wgrad [{cin:4d}->{cout:4d} @{res:2d}] {t * 1e3:8.1f} us  {flops / t / 1e9:7.1f} TF/s", flush=True)
