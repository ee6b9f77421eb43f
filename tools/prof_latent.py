"""Kernel-time breakdown of one eager training step of a latent config (bench.py --config celebahq|div2k), torch.profiler / CUPTI.
Usage: python tools/prof_latent.py celebahq|div2k [batch]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

name = sys.argv[1] if len(sys.argv) > 1 else "celebahq"
spec = bench.LATENT_CONFIGS[name]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
ldm, cfg = bench.build_from_yaml(spec["yaml"], dev)
B = int(sys.argv[2]) if len(sys.argv) > 2 else spec["batch"]
size = cfg["model"]["image_size"]
batch = {"image": 2 * torch.rand(B, 3, *size, device=dev) - 1}
if name == "div2k":
    d = ldm.first_stage_model.down_ratio
    batch["cond"] = 2 * torch.rand(B, 3, size[0] // d, size[1] // d, device=dev) - 1
ldm.train()
if name == "celebahq":
    from adm_b200.train import TrainStep
    step = TrainStep(ldm, lr=5e-5)

    def one(tag):
        with torch.no_grad():
            torch.cuda.nvtx.range_push("ae")
            z, *_ = ldm.get_input(batch)
            z = ldm.scale_factor * z
            torch.cuda.nvtx.range_pop()
        step.micro_step(z)
        step.optimizer_step()
else:
    params = [p for p in ldm.model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-5, fused=True)

    def one(tag):
        opt.zero_grad(set_to_none=True)
        loss, _ = ldm.training_step(batch)
        loss.backward()
        opt.step()
for _ in range(3):
    one("warm")
torch.cuda.synchronize()
# AE encode alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(3):
        ldm.get_input(batch)
    e1.record()
torch.cuda.synchronize()
print(f"{name} B={B}: frozen AE encode alone {e0.elapsed_time(e1) / 3:.2f} ms per batch")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    one("prof")
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        nm = ev.name.replace("void at::native::(anonymous namespace)::", "at::").replace("void at::native::", "at::")[:100]
        tot[nm][0] += 1
        tot[nm][1] += ev.device_time_total
allt = sum(v[1] for v in tot.values())
print(f"kernel time of one training step: {allt / 1000:.1f} ms, {sum(v[0] for v in tot.values())} kernels")
for nm, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:26]:
    print(f"{us / 1000:8.2f} ms {100 * us / allt:5.1f}%  n={c:5d}  avg {us / c:8.1f} us  {nm}")
