"""A few launches of the fused attention kernels at the CIFAR step's shapes (B = 128, C = 384, 6 heads; N = 256 and 64),
for `ncu -k regex:attn_`.  Usage: python tools/run_attn_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops

torch.manual_seed(0)
for side in (16, 8):
    qkv = (torch.randn(128, side, side, 3 * 384, device="cuda") * 0.8).bfloat16()
    da = torch.randn(128, side, side, 384, device="cuda").bfloat16()
    for _ in range(3):
        a, lse = ops.attention_fwd(qkv, 6)
        dqkv = ops.attention_bwd(da, qkv, lse, 6, a=a)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(20):
        a, lse = ops.attention_fwd(qkv, 6)
    e1.record()
    for _ in range(20):
        dqkv = ops.attention_bwd(da, qkv, lse, 6, a=a)
    e2.record()
    torch.cuda.synchronize()
    n = side * side
    fl = 4.0 * 384 * n * n * 128
    tf, tb = e0.elapsed_time(e1) / 20, e1.elapsed_time(e2) / 20
    print(f"N={n}: fwd {tf * 1000:.1f} us ({fl / tf / 1e9:.0f} TFLOP/s)  bwd {tb * 1000:.1f} us ({2.5 * fl / tb / 1e9:.0f} TFLOP/s, "
          f"5 products incl. the recomputed S)")
