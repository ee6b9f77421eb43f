"""Kernel timeline of graph-replayed training steps (or sampler forwards) through torch.profiler / CUPTI: per kernel
name count, total and average duration, plus GPU busy time against the wall time of the replay (idle = launch gaps).
Unlike ncu's launch list this is in-context (warm L2, back-to-back).  Usage: python tools/prof_step.py [train|sample] [B]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from adm_b200.train import TrainStep

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
# under torchrun (WORLD_SIZE > 1) this profiles the data-parallel step: the graph chain plus the NCCL all-reduces between
# its segments; every rank profiles itself, rank 0 prints
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
dpm = bench.build_model(dev)
dpm.train()
step = TrainStep(dpm)
x = 2 * torch.rand(B, 3, 32, 32, device=dev) - 1
REPS = 3
if mode == "train":
    step.capture(x)
    for _ in range(3):
        step.replay(x)
    run = lambda: step.replay(x)
else:
    dpm.eval()
    dpm.sampling_timesteps = 2
    dpm.sample(batch_size=B)
    run = lambda: dpm.sample(batch_size=B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(REPS):
        run()
    e1.record()
    torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / REPS
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0]
        tot[name][0] += 1
        tot[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
busy = sum(v[1] for v in tot.values()) / 1000.0 / REPS
if rank != 0:
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
print(f"{mode} B={B} ranks={world}: wall {wall:.2f} ms per run (under the profiler), kernels busy {busy:.2f} ms, "
      f"{sum(v[0] for v in tot.values()) // REPS} kernels per run")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 1000 / REPS:8.3f} ms {100 * us / 1000 / REPS / wall:5.1f}%  n={n // REPS:5d}  avg={us / n:8.1f} us  {name[:100]}")
# ADM_PROF_TIMELINE=path: every kernel of the LAST profiled run in start order — start (us from the first kernel), duration,
# gap to the end of the previous kernel on the same stream, stream, name — to see where a latency-bound chain waits
tl = os.environ.get("ADM_PROF_TIMELINE")
if tl:
    evs = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            dur = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            evs.append((ev.time_range.start, dur, getattr(ev, "device_resource_id", -1), ev.name.split("(")[0]))
    evs.sort()
    evs = evs[len(evs) - len(evs) // REPS:]
    t0 = evs[0][0]
    last_end = {}
    with open(tl, "w") as f:
        for st, dur, stream, name in evs:
            gap = st - last_end.get(stream, st)
            last_end[stream] = st + dur
            short = name.replace("void ", "").replace("adm::", "")[:60]
            f.write(f"{st - t0:10.1f} {dur:8.1f} gap {gap:7.1f} s{stream} {short}\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
