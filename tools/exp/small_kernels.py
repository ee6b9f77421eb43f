"""Isolated timings of the step's small streaming kernels (column sums, skip adds, resamples) at their largest shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit

for (hw, c) in [(16, 1152), (8, 1152), (32, 192), (16, 384)]:
    xs = [torch.randn(128, hw, hw, c, device="cuda").bfloat16() for _ in range(3)]
    out = torch.zeros(c, device="cuda")
    i = [0]
    def f():
        i[0] += 1
        ops.col_sums(xs[i[0] % 3], out)
    t = timeit(f)
    print(f"col_sums [128,{hw},{hw},{c}]: {t*1000:6.1f} us  {xs[0].numel()*2/t/1e9:5.2f} TB/s", flush=True)
for (hw, c) in [(32, 192), (16, 384)]:
    a = [torch.randn(128, hw, hw, c, device="cuda").bfloat16() for _ in range(4)]
    t = timeit(lambda: ops.add_bf16(a[0], a[1]))
    print(f"add_bf16 [128,{hw},{hw},{c}]: {t*1000:6.1f} us  {a[0].numel()*6/t/1e9:5.2f} TB/s (3 x 2 B/elem)", flush=True)
    t = timeit(lambda: ops.resample(a[2], 2))
    print(f"resample up [128,{hw},{hw},{c}]: {t*1000:6.1f} us  {a[0].numel()*10/t/1e9:5.2f} TB/s (2 + 8 B/elem)", flush=True)
    t = timeit(lambda: ops.resample(a[3], 1))
    print(f"resample down [128,{hw},{hw},{c}]: {t*1000:6.1f} us  {a[0].numel()*2.5/t/1e9:5.2f} TB/s (2 + 0.5 B/elem)", flush=True)
