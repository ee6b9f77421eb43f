"""Isolated timings of the relation-layer kernels (csrc/relation_ops.cu) at the DIV2K config's shapes (cond_unet on
128x128 latents, batch 8 / 16): achieved HBM bytes per second against the measured copy bandwidth.  Algorithmic bytes:
rel_gn_fwd 10 B/elem (x2 and the conv output read twice as bf16, one bf16 write), rel_gn_bwd 14 B/elem, bilinear_bwd
2 B/elem of dy (+ the 1/r^2 outputs), avgpool fwd/bwd 2 B/elem of the full-resolution side.
Usage: python tools/bench_relation.py [batch]        (also the target of `ncu --set full -k regex:rel_gn`)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops
from tools.bench_convs_lib import timeit as _timeit_ms


def timeit(fn):
    return _timeit_ms(fn) / 1e3  # seconds per call


B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
peak = 6551.7
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
print(f"# batch {B}; HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth)")
NB = 3  # rotate over buffers so that the 126 MB L2 does not serve the second pass of the NEXT launch
for (hw, c, hq, win) in [(128, 128, 16, 8), (64, 256, 16, 4), (32, 512, 16, 2), (16, 512, 16, 1)]:
    xs = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(NB)]
    ys = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(NB)]
    ds = [torch.randn(B, hw, hw, c, device="cuda").bfloat16() for _ in range(NB)]
    z = torch.randn(B, hq, hq, c, device="cuda")
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    n = xs[0].numel()
    _, stats = ops.rel_gn_fwd(xs[0], ys[0], z, gamma, beta, 8)
    i = [0]

    def fwd():
        i[0] += 1
        k = i[0] % NB
        ops.rel_gn_fwd(xs[k], ys[k], z, gamma, beta, 8)

    def bwd():
        i[0] += 1
        k = i[0] % NB
        ops.rel_gn_bwd(ds[k], xs[k], ys[k], gamma, stats, 8)

    def bil():
        i[0] += 1
        ops.bilinear_bwd(ds[i[0] % NB], (hq, hq))

    t = timeit(fwd)
    print(f"rel_gn_fwd   [{B},{hw},{hw},{c}] <- {hq}x{hq}: {t * 1e6:7.1f} us  {10 * n / t / 1e9:7.0f} GB/s  {10 * n / t / 1e9 / peak:.2f} of peak", flush=True)
    t = timeit(bwd)
    print(f"rel_gn_bwd   [{B},{hw},{hw},{c}]           : {t * 1e6:7.1f} us  {14 * n / t / 1e9:7.0f} GB/s  {14 * n / t / 1e9 / peak:.2f} of peak (+ the dgamma/dbeta reduce)", flush=True)
    if hq != hw:
        t = timeit(bil)
        print(f"bilinear_bwd [{B},{hw},{hw},{c}] -> {hq}x{hq}: {t * 1e6:7.1f} us  {2 * n / t / 1e9:7.0f} GB/s  {2 * n / t / 1e9 / peak:.2f} of peak", flush=True)
    if win > 1:
        def pf():
            i[0] += 1
            ops.avgpool_fwd(xs[i[0] % NB], (win, win))
        dyp = torch.randn(B, hw // win, hw // win, c, device="cuda").bfloat16()
        t = timeit(pf)
        print(f"avgpool_fwd  [{B},{hw},{hw},{c}] / {win}x{win}   : {t * 1e6:7.1f} us  {2 * n / t / 1e9:7.0f} GB/s  {2 * n / t / 1e9 / peak:.2f} of peak", flush=True)
        t = timeit(lambda: ops.avgpool_bwd(dyp, (B, hw, hw, c), (win, win)))
        print(f"avgpool_bwd  [{B},{hw},{hw},{c}] / {win}x{win}   : {t * 1e6:7.1f} us  {2 * n / t / 1e9:7.0f} GB/s  {2 * n / t / 1e9 / peak:.2f} of peak", flush=True)
    del xs, ys, ds
