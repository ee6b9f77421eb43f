"""Fixed cost of one GEMM-engine launch inside a CUDA graph: tiny problems (one tile, one k-iteration) vs epilogue-heavy
ones (many columns, tiny K), so that launch + prologue + teardown and the per-tile epilogue can be separated."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit
dev = "cuda"
def gemm(m, n, k):
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    c = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm_nt(a, b, out=c), iters=20)
    tiles = ((m + 127) // 128) * ((n + 191) // 192 if n > 192 else 1)
    print(f"gemm_nt M={m:6d} N={n:4d} K={k:5d}: {t*1000:7.2f} us  ({2.0*m*n*k/t/1e9:7.1f} TF/s)", flush=True)
for m, n, k in [(128, 64, 64), (128, 192, 64), (128 * 148, 192, 64), (128 * 148, 192, 384), (128 * 148, 192, 3456),
                (128 * 148 * 2, 192, 64), (128 * 148 * 2, 192, 384), (128 * 148 * 4, 192, 384), (32768, 1152, 384), (8192, 1152, 384),
                (8192, 384, 384), (2048, 384, 3456), (2048, 384, 384)]:
    gemm(m, n, k)
x = torch.randn(128, 4, 4, 384, device=dev).bfloat16()
g, b = torch.ones(384, device=dev), torch.zeros(384, device=dev)
t = timeit(lambda: ops.gn_forward(x, None, g, b, 32), iters=20)
print(f"gn_forward [128,4,4,384]: {t*1000:.2f} us")
e = torch.zeros(16, device=dev)
t = timeit(lambda: e.add_(1.0), iters=20)
print(f"torch tiny add_: {t*1000:.2f} us")
