#!/usr/bin/env python
"""Conditional sampling entry point: the role of /root/reference/sample_cond_ldm.py (whole-image path, :158-218):

    torchrun --nproc-per-node N scripts/sample_cond_ldm.py --cfg configs/super-resolution/div2k_cond_ddm_const_ldm.yaml \
        [--ckpt results/.../model-k.pt] [--num 16] [--out sr.pt]

Builds the model from the YAML, loads ``model-{k}.pt`` (EMA weights when ``sampler.use_ema``, :140-154) and runs
``LatentDiffusion.sample(cond=...)`` over the condition batches, sharded over the ranks with no communication.  On this path
the whole N-step latent loop replays as one CUDA graph and the condition encoder runs once per batch.  The
reference's sliding-window variants (``slide_sample*``, :220-332) are dataset glue around the same ``sample`` call and are
not reproduced here.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.train_uncond_ldm import build_model  # noqa: E402


def load_weights(model, ckpt_path, use_ema=True):
    """sample_cond_ldm.py:140-154: the EMA copy ('ema_model.*' inside data['ema']) or the online weights."""
    data = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    if use_ema and "ema" in data:
        sd = {k[len("ema_model."):]: v for k, v in data["ema"].items() if k.startswith("ema_model.")}
    else:
        sd = data["model"]
    return model.load_state_dict(sd)


def condition_batches(cfg, batch, num, down, seed):
    d = cfg["data"]
    if d.get("class_name", "synthetic") == "synthetic":
        g = torch.Generator().manual_seed(seed)
        h, w = d["image_size"][0] // down, d["image_size"][1] // down
        for i in range(0, num, batch):
            yield {"cond": 2 * torch.rand(min(batch, num - i), 3, h, w, generator=g) - 1}
        return
    from adm_b200.ddm.utils import construct_class_by_name
    ds = construct_class_by_name(**{k: v for k, v in d.items() if k not in ("batch_size", "num_workers")})
    seen = 0
    for b in torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=False, num_workers=d.get("num_workers", 0)):
        if seen >= num:
            return
        seen += b["cond"].shape[0]
        yield b


@torch.no_grad()
def sample_all(model, batches, device, rank=0, world=1):
    out = []
    for i, b in enumerate(batches):
        if i % world != rank:  # batches are dealt round-robin to the ranks
            continue
        cond = b["cond"].to(device)
        mask = b["ori_mask"].to(device) if "ori_mask" in b else None
        out.append(model.sample(batch_size=cond.shape[0], cond=cond, mask=mask))
    return torch.cat(out) if out else torch.empty(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--num", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cfg = yaml.safe_load(open(args.cfg))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    s = cfg.get("sampler", {})
    torch.manual_seed(1234)
    model = build_model(cfg, device).eval()
    ckpt = args.ckpt or s.get("ckpt_path")
    if ckpt and os.path.isfile(ckpt):
        print(load_weights(model, ckpt, s.get("use_ema", True)), flush=True)
    torch.manual_seed(99 + rank)  # per-rank start noise
    batch = args.batch or s.get("batch_size", cfg["data"]["batch_size"])
    num = args.num or s.get("sample_num", batch)
    imgs = sample_all(model, condition_batches(cfg, batch, num, model.first_stage_model.down_ratio, seed=7), device, rank, world)
    if args.out:
        torch.save(imgs.float().cpu(), f"{args.out}.rank{rank}" if world > 1 else args.out)
    print(f"rank {rank}: sampled {tuple(imgs.shape)} in [{float(imgs.min()):.3f}, {float(imgs.max()):.3f}]", flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
