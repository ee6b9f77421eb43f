"""Accounts for the GEMM-engine time of one training step: records every conv_fprop / conv_dgrad / conv_wgrad call of an
eager step (shapes), times each distinct call in isolation (graph of 10 launches, L2-warm), and prints count x time —
to be compared with the in-context CUPTI totals (tools/prof_step.py)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from adm_b200 import ops
from adm_b200.train import TrainStep
from tools.bench_convs_lib import timeit

dev = torch.device("cuda", 0)
dpm = bench.build_model(dev)
dpm.train()
step = TrainStep(dpm)
x = 2 * torch.rand(128, 3, 32, 32, device=dev) - 1
for _ in range(2):
    step.micro_step(x); step.optimizer_step()
torch.cuda.synchronize()
calls = collections.OrderedDict()
orig = {n: getattr(ops, n) for n in ("conv_fprop", "conv_dgrad", "conv_wgrad")}

def rec(name):
    def inner(*a, **k):
        if name == "conv_fprop":
            x1, wpk = a[0], a[1]; x2 = k.get("x2")
            key = (name, tuple(x1.shape), x2.shape[-1] if x2 is not None else 0, tuple(wpk.shape), k.get("residual") is not None, str(k.get("out_dtype", torch.bfloat16)))
        elif name == "conv_dgrad":
            dy, wpk = a[0], a[1]
            key = (name, tuple(dy.shape), 0, tuple(wpk.shape), k.get("residual") is not None, "")
        else:
            dy, x1 = a[0], a[1]; x2 = k.get("x2")
            key = (name, tuple(dy.shape), x2.shape[-1] if x2 is not None else 0, (x1.shape[-1], k.get("ntaps", 9)), False, "")
        calls.setdefault(key, [0, (a, k)])[0] += 1
        return orig[name](*a, **k)
    return inner

for n in orig:
    setattr(ops, n, rec(n))
step.micro_step(x); step.optimizer_step()
torch.cuda.synchronize()
for n in orig:
    setattr(ops, n, orig[n])
tot = collections.defaultdict(float)
rows = []
for key, (cnt, (a, k)) in calls.items():
    name = key[0]
    k = dict(k)
    if name == "conv_wgrad":
        k["out"] = torch.zeros_like(orig[name](*a, **{kk: vv for kk, vv in k.items() if kk != "out"}))
    t = timeit(lambda: orig[name](*a, **k))
    tot[name] += cnt * t
    rows.append((cnt * t, cnt, t, key))
for ct, cnt, t, key in sorted(rows, reverse=True)[:45]:
    print(f"{ct:7.3f} ms = {cnt:3d} x {t*1000:7.1f} us  {key}")
print({k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2), "ms; distinct", len(rows), "calls", sum(r[1] for r in rows))
