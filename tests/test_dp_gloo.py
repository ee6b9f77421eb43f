"""CPU, world_size 2 over gloo: replica synchronisation at construction and the overlapped bucketed gradient reduction of
TrainStep (host logic of the DP path).
The engine's backward is simulated by filling rank-dependent gradients and firing the gradient-ready hooks in
completion order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden.make_golden import TINY


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adm_b200.ddm.ddm_const import DDPM
    from adm_b200.train import TrainStep
    from adm_b200.unet.uncond_unet import EDMPrecond
    torch.manual_seed(100 + rank)  # replicas start from DIFFERENT weights: TrainStep must broadcast rank 0's
    kw = {k: v for k, v in TINY.items() if k not in ("img_resolution", "img_channels", "label_dim")}
    net = EDMPrecond(img_resolution=16, img_channels=3, sigma_data=1.0, model_type="DhariwalUNet", **kw)
    cfg = dict(image_size=[16, 16], sampling_timesteps=3)
    dpm = DDPM(model=net, cfg=cfg, **cfg)
    step = TrainStep(dpm, bucket_mb=1)  # ~6 M parameters -> many 1 MiB buckets
    assert step.world == world and len(step.buckets) > 8
    a = step.arena
    flats = [torch.empty_like(a.flat) for _ in range(world)]
    dist.all_gather(flats, a.flat)
    same_weights = all(torch.equal(f, flats[0]) for f in flats) and float(a.flat.abs().sum()) > 0
    # fake backward: rank r contributes (r + 1) * ramp; hooks fire block by block in completion order
    ramp = torch.arange(a.numel, dtype=torch.float32) % 97
    step._arm()
    a.grads.copy_((rank + 1) * ramp)
    launched_early = 0
    chunk = max(1, len(a.params) // 23)
    for i in range(0, len(a.params), chunk):
        step._on_grads(a.params[i:i + chunk])
        launched_early = step._next_bucket
    assert 0 < launched_early  # some buckets were reduced before "backward" finished
    step._allreduce()
    assert step._next_bucket == len(step.buckets)
    expect = sum(r + 1 for r in range(world)) * ramp
    ok = torch.equal(a.grads, expect)
    # parameters see the reduced gradients through their .grad views
    p = a.params[5]
    ok = ok and torch.equal(p.grad.flatten(), expect[a.offsets[5]:a.offsets[5] + p.numel()])
    out[rank] = bool(ok and same_weights)
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
