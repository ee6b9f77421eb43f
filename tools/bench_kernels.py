"""Micro-benchmarks of the HBM-bound kernels (CUDA events on the launching stream, inputs rotated over buffers larger
than L2).  Prints one JSON line per kernel with algorithmic bytes, time and GB/s against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from adm_b200 import ops

PK = bench.peaks()


def timeit(fn, iters=20, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, nbytes, ms, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / PK["hbm"], 3), "peak": PK["hbm"], "note": note}), flush=True)


def main():
    dev = "cuda"
    torch.manual_seed(0)
    # ---- K1/K2/K3 at a >= 256 MiB working set: [16384, 3, 32, 32] fp32 = 201 MB per tensor
    b = 16384
    nb = 2
    xs = [torch.rand(b, 3, 32, 32, device=dev) * 2 - 1 for _ in range(nb)]
    es = [torch.randn(b, 3, 32, 32, device=dev) for _ in range(nb)]
    cp = [torch.randn(b, 3, 32, 32, device=dev) for _ in range(nb)]
    ep = [torch.randn(b, 3, 32, 32, device=dev) for _ in range(nb)]
    t = torch.rand(b, device=dev) * 0.99 + 1e-4
    numel = xs[0].numel()
    ms = timeit(lambda i: ops.qsample(xs[i % nb], es[i % nb], t))
    report("K1 qsample fp32 [16384,3,32,32]", 12 * numel, ms, "read x0, eps; write x_t")
    ms = timeit(lambda i: ops.ddm_loss(cp[i % nb], ep[i % nb], xs[i % nb], es[i % nb], t, 1e-4, True))
    report("K2 ddm_loss fwd+bwd fp32", 24 * numel, ms, "read C^, eps^, x0, eps; write dC^, deps^")
    x64 = [x.double() for x in xs]
    ms = timeit(lambda i: ops.sampler_step(x64[i % nb], cp[i % nb], ep[i % nb], 0.7, 0.6))
    report("K3 sampler_step fp64 state", 24 * numel, ms, "read x (f64), C^, eps^ (f32); write x' (f64)")
    xf = [x.float() for x in xs]
    ms = timeit(lambda i: ops.sampler_step(xf[i % nb], cp[i % nb], ep[i % nb], 0.7, 0.6))
    report("K3 sampler_step fp32 state", 16 * numel, ms, "read x, C^, eps^; write x'")
    del xs, es, cp, ep, x64, xf
    # ---- config-shape latency of the same kernels (B = 128: 4.7 MB tensors, L2 resident, launch bound)
    b = 128
    x0 = torch.rand(b, 3, 32, 32, device=dev)
    e0 = torch.randn_like(x0)
    t = torch.rand(b, device=dev) * 0.99 + 1e-4
    ms = timeit(lambda i: ops.qsample(x0, e0, t), iters=200)
    report("K1 qsample @B=128 (latency)", 12 * x0.numel(), ms, "L2 resident")
    ms = timeit(lambda i: ops.ddm_loss(x0, e0, x0, e0, t, 1e-4, True), iters=200)
    report("K2 ddm_loss @B=128 (latency)", 24 * x0.numel(), ms, "L2 resident")
    # ---- GroupNorm family at the CIFAR shapes, rotating over > L2 worth of buffers
    for (n, hw, c) in [(128, 32, 192), (128, 16, 384), (128, 8, 384)]:
        nbuf = max(2, int(300e6 // (n * hw * hw * c * 2)) + 1)
        xs = [torch.randn(n, hw, hw, c, device=dev).bfloat16() for _ in range(nbuf)]
        dys = [torch.randn(n, hw, hw, c, device=dev).bfloat16() for _ in range(nbuf)]
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        params = torch.randn(n, 2 * c, device=dev) * 0.1
        g = min(32, c // 4)
        elems = n * hw * hw * c
        ms = timeit(lambda i: ops.gn_stats(xs[i % nbuf], None, gamma, beta, g, 1e-5, params=params))
        report(f"gn_stats [{n},{hw},{hw},{c}]", 2 * elems, ms, "read x")
        coef = ops.gn_stats(xs[0], None, gamma, beta, g, 1e-5, params=params)
        ms = timeit(lambda i: ops.gn_apply(xs[i % nbuf], None, coef, act=True, drop_p=0.1, seed=i))
        report(f"gn_apply+silu+dropout [{n},{hw},{hw},{c}]", 4 * elems, ms, "read x; write y")
        ms = timeit(lambda i: ops.gn_forward(xs[i % nbuf], None, gamma, beta, g, 1e-5, params=params, act=True,
                                             drop_p=0.1, seed=i))
        report(f"gn_forward stats+apply+silu+dropout [{n},{hw},{hw},{c}]", 6 * elems, ms,
               "read x (stats), read x again (L2), write y")
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        dp = torch.zeros(n, 2 * c, device=dev)
        ms = timeit(lambda i: ops.gn_bwd(dys[i % nbuf], xs[i % nbuf], None, coef, gamma, beta, g, params=params, act=True,
                                         drop_p=0.1, seed=i, dgamma=dg, dbeta=db, dparams=dp, add=dys[(i + 1) % nbuf]))
        report(f"gn_bwd (reduce+apply+skip add) [{n},{hw},{hw},{c}]", 12 * elems, ms,
               "reduce: read dy, x; apply: read dy, x, add; write dx")
        del xs, dys
    # ---- optimizer over a 216 M element arena
    n = 216_141_136
    p, g_, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    sq = torch.zeros(1, device=dev)
    ms = timeit(lambda i: ops.adamw(p, g_, m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-4, 5, 1.0, 1.0, sq), iters=10)
    report("K10 clip+AdamW 216M params", 28 * n, ms, "read p, g, m, v; write p, m, v")
    ms = timeit(lambda i: ops.sq_norm(g_, sq), iters=10)
    report("grad sq-norm 216M", 4 * n, ms, "read g")


if __name__ == "__main__":
    main()
