"""Which operand's TMA traffic does the conv main loop wait for?  Run once per ADM_GEMM_DEBUG value (read at first use):
0 = normal, 4 = only the weight tile is loaded, 8 = only the pixel tile is loaded (results are garbage, timing is the point)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit

N = 128
for cin, cout, res in [(384, 384, 16), (192, 192, 32), (768, 384, 16), (256, 256, 16), (128, 128, 32)]:
    flops = 2.0 * N * res * res * cin * cout * 9
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 60)
    x = torch.randn(N, res, res, cin, device="cuda").bfloat16()
    out = torch.empty(N, res, res, cout, device="cuda", dtype=torch.bfloat16)
    t = timeit(lambda: ops.conv_fprop(x, w, out=out))
    print(f"debug={os.environ.get('ADM_GEMM_DEBUG', '0')} [{cin:4d}->{cout:4d} @{res:2d}] {t * 1e3:8.1f} us  {flops / t / 1e9:7.1f} TF/s", flush=True)
