"""Isolated timings for the GroupNorm-from-epilogue-statistics path: conv with / without the statistics epilogue, the
cluster kernel (stats + apply) against finalize + streaming apply.  CUDA events, rotating buffers larger than L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops


def timeit(fn, iters=40, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000  # us


def main():
    dev = "cuda"
    n = 128
    for (hw, cin, cout, k) in [(16, 384, 384, 3), (32, 192, 192, 3), (8, 384, 384, 3), (4, 384, 384, 3), (16, 384, 384, 1)]:
        nbuf = max(2, int(300e6 // (n * hw * hw * (cin + cout) * 2)) + 1)
        xs = [torch.randn(n, hw, hw, cin, device=dev).bfloat16() for _ in range(nbuf)]
        w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device=dev) / (k * cin ** 0.5))
        bias = torch.zeros(cout, device=dev)
        t0 = timeit(lambda i: ops.conv_fprop(xs[i % nbuf], w, bias=bias))
        t1 = timeit(lambda i: ops.conv_fprop(xs[i % nbuf], w, bias=bias, stats=True))
        print(f"conv {cin}->{cout} k{k} @{hw}: plain {t0:7.1f} us   +stats epilogue {t1:7.1f} us  ({100 * (t1 / t0 - 1):+.1f} %)",
              flush=True)
    for (hw, c1, c2) in [(32, 192, 0), (16, 384, 0), (16, 384, 384), (32, 384, 192), (8, 384, 0), (8, 384, 384), (4, 384, 0)]:
        c = c1 + c2
        nbuf = max(2, int(300e6 // (n * hw * hw * c * 4)) + 1)
        x1 = [torch.randn(n, hw, hw, c1, device=dev).bfloat16() for _ in range(nbuf)]
        x2 = [torch.randn(n, hw, hw, c2, device=dev).bfloat16() for _ in range(nbuf)] if c2 else None
        slots = max(1, hw * hw // 32)
        st1 = torch.randn(n + 1, slots, c1, 2, device=dev).abs() + 1
        st1[..., 1] += st1[..., 0] ** 2
        st2 = None
        if c2:
            st2 = torch.randn(n + 1, slots, c2, 2, device=dev).abs() + 1
            st2[..., 1] += st2[..., 0] ** 2
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        params = 0.1 * torch.randn(n, 2 * c, device=dev)
        g = min(32, c // 4)
        tf = timeit(lambda i: ops.gn_forward(x1[i % nbuf], x2[i % nbuf] if c2 else None, gamma, beta, g, 1e-5, params=params,
                                             act=True, drop_p=0.1, seed=i))
        ts = timeit(lambda i: ops.gn_forward_stats(x1[i % nbuf], st1, x2[i % nbuf] if c2 else None, st2, gamma, beta, g, 1e-5,
                                                   params=params, act=True, drop_p=0.1, seed=i))
        coef = torch.randn(n, c, 4, device=dev)
        ta = timeit(lambda i: ops.gn_apply(x1[i % nbuf], x2[i % nbuf] if c2 else None, coef, act=True, drop_p=0.1, seed=i))
        mb = n * hw * hw * c * 4 / 1e6
        print(f"GN [{n},{hw},{hw},{c1}+{c2}] ({mb:6.1f} MB r+w): cluster stats+apply {tf:6.1f} us | finalize+apply {ts:6.1f} us "
              f"| apply alone {ta:6.1f} us = {mb / ta:.2f} TB/s", flush=True)


if __name__ == "__main__":
    main()
