"""In-context time per op category for one eager training step (or one sampler forward): every `adm_b200.ops` wrapper is
bracketed by CUDA events on the launching stream, so the numbers are warm-L2, back-to-back timings (unlike ncu's
serialised cold-cache launch list).  Usage: python tools/time_ops.py [train|sample] [batch]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from adm_b200 import ops
from adm_b200.train import TrainStep

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
dpm = bench.build_model(dev)
dpm.train()
step = TrainStep(dpm)
x = 2 * torch.rand(B, 3, 32, 32, device=dev) - 1
for _ in range(3):
    step.micro_step(x)
    step.optimizer_step()
torch.cuda.synchronize()

records = []
depth = [0]


def wrap(name, fn):
    def inner(*a, **k):
        if depth[0] > 0:  # nested wrapper (attention_fwd -> softmax_fwd): account to the outer op
            return fn(*a, **k)
        depth[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            out = fn(*a, **k)
        finally:
            depth[0] -= 1
        e1.record()
        tag = name
        if name in ("conv_fprop", "conv_dgrad", "conv_wgrad"):
            xx = a[0]
            taps = a[1].shape[1] if name != "conv_wgrad" else k.get("ntaps", 9)
            tag = f"{name} k{1 if taps == 1 else 3} @{xx.shape[1]}"
        elif name in ("gn_stats", "gn_apply", "gn_bwd"):
            xx = a[1] if name == "gn_bwd" else a[0]
            tag = f"{name} @{xx.shape[1]}"
        records.append((tag, e0, e1))
        return out
    return inner


skip = {"pad64", "check", "set_seed_counter"}
for name in dir(ops):
    fn = getattr(ops, name)
    if callable(fn) and not name.startswith("_") and name not in skip and getattr(fn, "__module__", "") == ops.__name__:
        setattr(ops, name, wrap(name, fn))

t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
if mode == "train":
    step.micro_step(x)
    step.optimizer_step()
else:
    dpm.eval()
    dpm.sampling_timesteps = 2
    dpm.sample(batch_size=B)
t1.record()
torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for tag, e0, e1 in records:
    tot[tag][0] += 1
    tot[tag][1] += e0.elapsed_time(e1)
wall = t0.elapsed_time(t1)
acc = sum(v[1] for v in tot.values())
print(f"{mode} B={B}: wall {wall:.2f} ms, inside ops {acc:.2f} ms ({len(records)} calls)")
for tag, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100 * ms / wall:5.1f}%  n={n:4d}  avg={1000 * ms / n:8.1f} us  {tag}")
