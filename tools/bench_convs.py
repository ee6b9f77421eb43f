"""Times every distinct conv shape of the CIFAR UNet (batch 128) in isolation: fprop, dgrad (MN-major view of the fprop-packed
weights), dgrad_T (fprop kernels on the transposed weight shadow: what the training step runs), wgrad -> TFLOP/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adm_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
SHAPES = [  # (count, cin, cout, res, k)
    (15, 384, 384, 16, 3), (14, 192, 192, 32, 3), (6, 768, 384, 16, 3), (6, 384, 192, 32, 3), (2, 384, 384, 32, 3),
    (2, 576, 192, 32, 3), (17, 384, 384, 8, 3), (8, 768, 384, 8, 3), (23, 384, 384, 4, 3), (8, 768, 384, 4, 3),
    (11, 384, 1152, 16, 1), (11, 384, 1152, 8, 1), (11, 384, 384, 16, 1), (6, 768, 384, 16, 1), (6, 384, 192, 32, 1),
    (1, 192, 384, 16, 1),
]

SHAPES = SHAPES[:int(os.environ.get("ADM_BENCH_SHAPES", len(SHAPES)))]  # first k shapes only

from tools.bench_convs_lib import timeit  # noqa: E402


tot = {"fprop": 0.0, "dgrad": 0.0, "dgrad_T": 0.0, "wgrad": 0.0}
for cnt, cin, cout, res, k in SHAPES:
    x = torch.randn(N, res, res, cin, device="cuda").bfloat16()
    dy = torch.randn(N, res, res, cout, device="cuda").bfloat16()
    w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") / (k * cin ** 0.5))
    out = torch.empty(N, res, res, cout, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(N, res, res, cin, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(cout, k * k, cin, device="cuda")
    flops = 2.0 * N * res * res * cin * cout * k * k
    r = {}
    r["fprop"] = timeit(lambda: ops.conv_fprop(x, w, out=out))
    r["dgrad"] = timeit(lambda: ops.conv_dgrad(dy, w, out=dx))
    # the training step's data gradient: the fprop kernels on the transposed, tap-mirrored weight shadow (dgrad_T)
    wt = torch.empty(cin, k * k, cout, device="cuda", dtype=torch.bfloat16)
    ops.transpose_weight_tiles(w, wt, ops.weight_transpose_tiles(0, cout, k * k, cin).cuda())
    r["dgrad_T"] = timeit(lambda: ops.conv_fprop(dy, wt, out=dx))
    r["wgrad"] = timeit(lambda: ops.conv_wgrad(dy, x, ntaps=k * k, out=dw))
    for kk in tot:
        tot[kk] += cnt * r[kk]
    print(f"{cnt:3d} x [{cin:4d}->{cout:4d} @{res:2d} k{k}] GFLOP {flops/1e9:7.1f} | " +
          " | ".join(f"{kk} {r[kk]*1000:7.1f} us {flops/r[kk]/1e9:7.1f} TF/s" for kk in ("fprop", "dgrad", "dgrad_T", "wgrad")), flush=True)
print("weighted totals (ms):", {k: round(v, 2) for k, v in tot.items()})
